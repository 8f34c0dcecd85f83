// ttb_mma.cuh -- level kernels for large alphabets (amino acids, q = 20..22) on the fp64 tensor pipe.
//
// Why: at q = 20 the one-thread-per-pattern kernels spend per child 800 DFMAs AND as many broadcast shared-memory loads
// of exp(Qt) (every P[i][j] is used for exactly one FMA per thread), so the LSU, not the fp64 pipe or HBM, bounds them
// (ncu: fp64 pipe 32 %, profiles/r3_pre_level_cfg4_ncu.txt), and 1 000 patterns expose only 1 000 threads per node.
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) has the same peak as DFMA on B200 (measured 63.6 vs 57 FMA/clk/SM,
// tools/probe/dmma_probe.cu) but takes its operands as register fragments: a fragment of P is loaded ONCE per
// (warp, child) and reused for all the warp's patterns, 256 FMAs per instruction.
//
// Mapping.  Patterns are the M dimension (8 per mma), states are N and K, padded 20..22 -> 24 = 3 n-tiles = 6 k-steps.
// The accumulator layout of the mma IS the state-over-lanes split: lane (g = lane/4, c = lane%4) of a warp holds, for
// pattern 8*mt + g of m-tile mt, the SIX states  j(nt, r) = 8*nt + 2*c + r  (nt = 0..2, r = 0..1)  -- register slot
// k = 2*nt + r.  Everything elementwise (products over children, the outside message Mp/U, the Hadamard product with
// the subtree profile, normalisers, argmax) happens in this layout with two xor-shuffles per reduction.  The K index of
// a product may be walked in any order, so k-step kap = 2*nt + r takes register slot kap of the same layout as its A
// operand -- no layout conversion between the two chained products of the preorder:
//    U[pat][j]   = sum_i S[pat][i] P[i][j]     A = S (slot kap),  B1[kap][nt'] : lane (g,c) holds P[8nt+2c+r][8nt'+g]
//    msg[pat][i] = sum_j O[pat][j] P[i][j]     A = O (slot kap),  B2[kap][nt'] : lane (g,c) holds P[8nt'+g][8nt+2c+r]
// The B fragments of every branch are laid out in this order by expqt_frag_kernel (pairs of fragments per lane,
// zero-padded) so a stage receives them with one bulk copy and a lane reads its elements conflict-free.
//
// Reference semantics are those of post_level_kernel / pre_level_kernel (treeanc.py:857-930); summation order differs
// (tolerances: log-LH 1e-9 relative, profiles 1e-6).  Single model, double storage, no masks, no joint pass: those keep
// the one-thread-per-pattern kernels.
#pragma once
#include "ttb_kernels.cuh"

#define TTB_MMA_NT 3                         // n-tiles of 8 states
// Register slots per lane and pattern = k-steps of a product.  Slot k of lane column c holds state
//   k < 4: 8*(k/2) + 2c + k%2      k = 4: 16 + c      k = 5: 20 + c
// i.e. the accumulator columns 17, 19, 21, 23 carry the states 20..23.  For q = 20 slot 5 is pure padding in EVERY lane,
// so both products run 5 k-steps (K = 20 exactly, 15 instead of 18 DMMAs) and every elementwise loop has 5 slots.
template <int Q>
struct MmaQ {
  static constexpr int KS = (Q + 3) / 4;            // 5 for q = 20, 6 for q = 21..24
  static constexpr int NF = KS * TTB_MMA_NT;        // fragments per product
  static constexpr int NP = (NF + 1) / 2;           // fragment pairs (one 16-byte load per lane)
  static constexpr int PFQ = NP * 64;               // doubles per product in the fragment-ordered exp(Qt)
  static_assert(2 * PFQ <= TTB_PF_STRIDE, "fragment-ordered exp(Qt): two products");
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double quad_sum(double x) {   // sum over the four lanes that share a pattern
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  return x;
}
__device__ __forceinline__ double quad_max(double x) {
  x = fmax(x, __shfl_xor_sync(0xffffffffu, x, 1));
  x = fmax(x, __shfl_xor_sync(0xffffffffu, x, 2));
  return x;
}
// state held in register slot k by lane column c
__device__ __forceinline__ int mma_state(int k, int c) { return k < 4 ? 8 * (k >> 1) + 2 * c + (k & 1) : 16 + 4 * (k - 4) + c; }

// A1 for the large alphabets in one kernel: exp(Qt) of a branch (same arithmetic and summation order as expqt_kernel, so
// P is bit-identical) written in both layouts -- P[i][j] for the tip tables, the fetch / branch kernels, and the fragment
// order above for the level kernels.  One warp per branch: the Q exponentials are taken once per branch (expqt_kernel
// takes them once per row: 20x), e_k Vinv[k][j] is staged in shared memory, and both outputs are coalesced stores.
// cfg4: expqt_kernel 45 us + pfrag_kernel 37 us -> one launch.
#define TTB_EXPQT_WARPS 4
#define TTB_EXPQT_NODES 2     // branches per warp (more would starve the SMs of warps at 10 000 branches): amortises the block's fragment-order index table
template <int Q>
__global__ void __launch_bounds__(TTB_EXPQT_WARPS * 32) expqt_frag_kernel(TtbDev p, double* __restrict__ Pf) {
  using MQ = MmaQ<Q>;
  __shared__ double sW[TTB_EXPQT_WARPS][Q * Q];   // e_k * Vinv[k][j]
  __shared__ double sP[TTB_EXPQT_WARPS][Q * Q];
  __shared__ double sE[TTB_EXPQT_WARPS][32];
  __shared__ double sV[Q * Q];
  __shared__ short sMap[2 * MQ::PFQ];             // fragment-order position -> i*Q + j of the source entry, -1 = padding
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = threadIdx.x; r < Q * Q; r += blockDim.x) sV[r] = p.v[r];
  for (int rem2 = threadIdx.x; rem2 < 2 * MQ::PFQ; rem2 += blockDim.x) {
    // rem2 = which * PFQ + ((f/2)*32 + lane)*2 + f%2
    const int which = rem2 / MQ::PFQ, rem = rem2 % MQ::PFQ;
    const int f = (rem >> 6) * 2 + (rem & 1), fl = (rem >> 1) & 31, fg = fl >> 2, fc = fl & 3;
    int src = -1;
    if (f < MQ::NF) {
      const int kap = f / TTB_MMA_NT, ntp = f % TTB_MMA_NT;
      const int kidx = mma_state(kap, fc), nidx = mma_state(2 * ntp + (fg & 1), fg >> 1);   // contracted state / state of accumulator column 8*ntp + fg
      const int i = which ? nidx : kidx, j = which ? kidx : nidx;
      if (i < Q && j < Q) src = i * Q + j;
    }
    sMap[rem2] = (short)src;
  }
  __syncthreads();
  const int node0 = (blockIdx.x * TTB_EXPQT_WARPS + warp) * TTB_EXPQT_NODES;
  for (int node = node0; node < min(node0 + TTB_EXPQT_NODES, p.n_nodes); ++node) {
    const double mt = p.mu[0] * p.t[node];
    if (lane < Q) sE[warp][lane] = exp(mt * p.eig[lane]);
    __syncwarp();
    for (int r = lane; r < Q * Q; r += 32) sW[warp][r] = sE[warp][r / Q] * p.vinv[r];
    __syncwarp();
    for (int r = lane; r < Q * Q; r += 32) {
      const int i = r / Q, j = r % Q;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < Q; ++k) acc = fma(sV[i * Q + k], sW[warp][k * Q + j], acc);
      acc = fmax(0.0, acc);
      sP[warp][r] = acc;
      p.P[(size_t)node * p.pq + r] = acc;
    }
    __syncwarp();
    double* __restrict__ out = Pf + (size_t)node * TTB_PF_STRIDE;
    for (int rem2 = lane; rem2 < 2 * MQ::PFQ; rem2 += 32) {
      const int src = sMap[rem2];
      out[rem2] = src >= 0 ? sP[warp][src] : 0.0;
    }
    __syncwarp();
  }
}

// C[mt][.] = A[mt][.] x B for the MT m-tiles of a warp; frag = this lane's column of the fragment pairs of one product.
// Accumulator column pair (2c, 2c+1) of n-tile ntp = register slots 2*ntp, 2*ntp + 1 (slot 5 is dropped when KS = 5).
template <int MT, int KS>
__device__ __forceinline__ void mma_product(const double (&A)[MT][KS], const double2* __restrict__ frag, double (&C)[MT][KS]) {
  double acc[MT][2 * TTB_MMA_NT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int k = 0; k < 2 * TTB_MMA_NT; ++k) acc[mt][k] = 0.0;
#pragma unroll
  for (int pp = 0; pp < (KS * TTB_MMA_NT + 1) / 2; ++pp) {
    const double2 b = frag[pp * 32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int f = 2 * pp + h, kap = f / TTB_MMA_NT, ntp = f % TTB_MMA_NT;
      if (f < KS * TTB_MMA_NT) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][2 * ntp], acc[mt][2 * ntp + 1], A[mt][kap], h ? b.y : b.x);
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int k = 0; k < KS; ++k) C[mt][k] = acc[mt][k];
}

// sum of a lane's slots (balanced tree: short dependency chain)
template <int KS>
__device__ __forceinline__ double slot_sum(const double (&x)[KS]) {
  double s = (x[0] + x[1]) + (x[2] + x[3]);
  if (KS == 5) return s + x[4];
  return s + (x[4] + x[KS - 1]);
}

// Measurement only: block timelines (see TtbDev::trace).  Slot = header {grid, block, smid, globaltimer, kernel id, chunks} +
// 32 clock64 events per warp.
#define TTB_TRACE_SLOTS 4096
#define TTB_TRACE_SLOT_WORDS (8 + 17 * 32)
__device__ __forceinline__ unsigned long long* trace_begin(const TtbDev& p, uint32_t* share, int kernel_id, int n_chunks) {
  if (!p.trace) return nullptr;   // uniform
  if (threadIdx.x == 0) {
    const unsigned long long s = (p.trace[1] == 0 || p.trace[1] == gridDim.x) ? atomicAdd(p.trace, 1ull) : ~0ull;   // trace[1]: grid-size filter
    *share = s < TTB_TRACE_SLOTS ? (uint32_t)s : 0xffffffffu;
    if (s < TTB_TRACE_SLOTS) {
      unsigned long long* b = p.trace + 16 + s * TTB_TRACE_SLOT_WORDS;
      unsigned int smid;
      unsigned long long gt;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      b[0] = gridDim.x; b[1] = blockIdx.x; b[2] = smid; b[3] = gt; b[4] = kernel_id; b[5] = n_chunks; b[6] = clock64();
    }
  }
  __syncthreads();
  const uint32_t s = *share;
  return s == 0xffffffffu ? nullptr : p.trace + 16 + (size_t)s * TTB_TRACE_SLOT_WORDS + 8 + (threadIdx.x >> 5) * 32;
}
__device__ __forceinline__ void trace_ev(unsigned long long* tr, int e) {
  if (tr && (threadIdx.x & 31) == 0 && e < 32) tr[e] = clock64();
}

template <int Q, int NW>
struct MmaCfg {
  static_assert(Q > 16 && Q <= 24, "three n-tiles");
  static_assert(16 % NW == 0, "whole m-tiles per warp");
  static constexpr int MT = TTB_TILE / 8 / NW;        // m-tiles (8 patterns) per warp
  static constexpr int THREADS = NW * 32;             // no dedicated producer warp (see below)
  // Register file: an SM sub-partition holds 16 384 registers and the warps of a block are dealt round-robin to the
  // four sub-partitions.  NW = 8 at <= 128 registers -> two blocks per SM (4 warps x 4 096 per sub-partition); a ninth
  // (producer) warp would put three warps of each block on one sub-partition and only one block would fit -- measured:
  // launch__occupancy_limit_registers = 1 at 9 x 32 x 112.  So warp 0 issues the bulk copies itself.
  static constexpr int MAXREG = NW == 16 ? 128 : (NW == 8 ? 128 : 255);   // NW = 16: one block per SM
  static constexpr int STAGES = 2;
  using PipeT = Pipe<Q, STAGES, NW * 32>;
};

// Pipeline without a producer warp: warp 0 issues the bulk copies of chunk u + STAGES - 1 when it STARTS chunk u (the
// stage it refills was released by every warp during chunk u - 1, so its wait on the empty barrier is almost always
// already satisfied; the copy then has a whole chunk of arithmetic to land).  Two independent blocks per SM keep the
// tensor pipe busy while the other one waits, divides or stores (block timelines: tools/trace_view.py).
template <typename PipeT>
struct SelfFed {
  typename PipeT::Cursor pc;   // producer cursor (warp 0 only)
  int issued = 0;
};

// ---------------------------------------------------------------------------------------
// Postorder level (A3-A5), see post_level_kernel.  Block = (run of nodes, 128-pattern tile), NW pattern warps,
// one child per chunk.
// ---------------------------------------------------------------------------------------
template <int Q, int NW>
__global__ void __launch_bounds__(MmaCfg<Q, NW>::THREADS) __maxnreg__((MmaCfg<Q, NW>::MAXREG)) post_level_mma_kernel(TtbDev p, const TtbChunk* __restrict__ chunks,
                                                                                 const int* __restrict__ group_ptr, int tiles, int fbase) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Cfg = MmaCfg<Q, NW>;
  using PipeT = typename Cfg::PipeT;
  constexpr int MT = Cfg::MT;
  constexpr int KS = MmaQ<Q>::KS;
  constexpr int PFQ = MmaQ<Q>::PFQ;
  constexpr uint32_t MSG_BYTES = Q * TTB_TILE * 8;
  constexpr int PFD = PFQ;   // only the first product's fragments
  PipeT pipe(smem_raw, Q, PFD, p.tu_stride);
  const int g_ = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int k0 = group_ptr[g_], k1 = group_ptr[g_ + 1];
  const int n_chunks = k1 - k0;
  const long long a0 = (long long)tile * TTB_TILE;
  const int cols = (int)min((long long)TTB_TILE, p.ld - a0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  if (tid == 0) pipe.init();
  __syncthreads();

  SelfFed<PipeT> feed;
  Chunk nextc = load_chunk_global(chunks + k0);   // warp 0: descriptor of the next chunk to issue
  auto issue_next = [&]() {   // warp 0
    const int u = feed.issued;
    if (u >= n_chunks) return;
    const Chunk c = nextc;
    nextc = load_chunk_global(chunks + k0 + min(u + 1, n_chunks - 1));
    const int s = feed.pc.s;
    pipe.producer_acquire(feed.pc, u);
    uint64_t* bar = pipe.full + s;
    const int src = c.src0;
    const bool skip = (p.dbg & 2) && src >= 0;
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, skip ? 32u : 32u + (src >= 0 ? MSG_BYTES + PFD * 8u : (uint32_t)(cols + p.tu_stride * 8)));
      tma_load_1d((void*)pipe.desc(s), chunks + k0 + u, 32, bar);
    }
    __syncwarp();
    if (skip) {
    } else if (src >= 0) {
      if (lane == 1) tma_load_1d(pipe.rows(s), p.S + msg_off<Q>(p, src, a0), MSG_BYTES, bar);
      if (lane == 2) tma_load_1d(pipe.P(s), p.Pf + (size_t)c.cnode0 * TTB_PF_STRIDE, PFD * 8, bar);
    } else {
      const int row = -1 - src;
      if (lane == 1) tma_load_1d(pipe.codes(s), p.codes + (size_t)row * p.ld + a0, cols, bar);
      if (lane == 2) tma_load_1d(pipe.TU(s), p.TU + (size_t)row * p.tu_stride, p.tu_stride * 8, bar);
    }
    feed.pc.advance();
    ++feed.issued;
  };
  pdl_wait();   // everything the bulk copies read was written by earlier levels
  if (warp == 0)
    for (int i = 0; i < Cfg::STAGES - 1; ++i) issue_next();

  typename PipeT::Cursor cur;
  const int g = lane >> 2, c4 = lane & 3;
  const int pat0 = warp * (MT * 8) + g;   // pattern (within the tile) of m-tile 0
  double X[MT][KS];
  double Facc[MT], Zprod[MT];
  int scale[MT];
  int seen = 0;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) { Facc[mt] = 0.0; Zprod[mt] = 1.0; scale[mt] = 0; }
  for (int u = 0; u < n_chunks; ++u) {
    const int s = cur.s;
    if (warp == 0) issue_next();   // refills the stage of chunk u - 1
    pipe.consumer_wait(cur);
    const Chunk c = load_chunk_smem(pipe.desc(s));
    if (c.flags & 1) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int k = 0; k < KS; ++k) X[mt][k] = 1.0;
        scale[mt] = 0;
      }
      seen = 0;
    }
    if (p.dbg & 1) {
      pipe.consumer_release(cur);
      cur.advance();
      continue;
    }
    double U[MT][KS];
    if (c.src0 >= 0) {
      double A[MT][KS];
      const double* rows = pipe.rows(s) + pat0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int i = mma_state(k, c4);
          A[mt][k] = (i < Q) ? rows[i * TTB_TILE + mt * 8] : 0.0;
        }
      mma_product<MT, KS>(A, reinterpret_cast<const double2*>(pipe.P(s)) + lane, U);
    } else {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int code = (a0 + pat0 + mt * 8 < p.Lp) ? pipe.codes(s)[pat0 + mt * 8] : 0;   // past the alignment: any valid row
        const double* tu = pipe.TU(s) + code * Q;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int j = mma_state(k, c4);
          U[mt][k] = (j < Q) ? tu[j] : 0.0;
        }
      }
    }
    pipe.consumer_release(cur);
    cur.advance();
    ++seen;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int k = 0; k < KS; ++k) X[mt][k] *= U[mt][k];
      if (seen > 2) {   // polytomy: keep the running product in range (exact scaling)
        double mx = X[mt][0];
#pragma unroll
        for (int k = 1; k < KS; ++k) mx = fmax(mx, X[mt][k]);
        mx = quad_max(mx);
        if (mx < 0x1p-256 && mx > 0.0) {
#pragma unroll
          for (int k = 0; k < KS; ++k) X[mt][k] *= 0x1p+256;
          ++scale[mt];
        }
      }
    }
    if (c.flags & 2) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // padded states hold 0 after the first child
        const double Z = quad_sum(slot_sum<KS>(X[mt]));
        const double inv = fast_rcp(Z);
        const long long a = a0 + pat0 + mt * 8;
        if (a < p.Lp) {
          double* __restrict__ so = p.S + msg_off<Q>(p, c.out, a);
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const int j = mma_state(k, c4);
            if (j < Q) so[j * TTB_TILE] = X[mt][k] * inv;
          }
        }
        if (scale[mt]) Facc[mt] -= scale[mt] * (256.0 * 0.693147180559945309417232121458);
        if (Z < 1e-150 || Z > 1e150) {
          Facc[mt] += log(Z);
        } else {
          Zprod[mt] *= Z;
          if (Zprod[mt] < 1e-150 || Zprod[mt] > 1e150) {
            Facc[mt] += log(Zprod[mt]);
            Zprod[mt] = 1.0;
          }
        }
      }
    }
  }
  if (c4 == 0) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const long long a = a0 + pat0 + mt * 8;
      if (a < p.Lp) p.Fpart[(size_t)(fbase + g_) * p.ld + a] = Facc[mt] + log(Zprod[mt]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Preorder level (A6-A7), see pre_level_kernel.  Stage rows: [0, Q) parent profile (first chunk of a parent),
// [Q, 2Q) the child's subtree profile.
// ---------------------------------------------------------------------------------------
template <int Q, int NW, bool TIPS>
__global__ void __launch_bounds__(MmaCfg<Q, NW>::THREADS) __maxnreg__((MmaCfg<Q, NW>::MAXREG)) pre_level_mma_kernel(TtbDev p, const TtbChunk* __restrict__ chunks,
                                                                                const int* __restrict__ group_ptr, int tiles, int count_diff) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Cfg = MmaCfg<Q, NW>;
  using PipeT = typename Cfg::PipeT;
  constexpr int MT = Cfg::MT;
  constexpr int KS = MmaQ<Q>::KS;
  constexpr int PFQ = MmaQ<Q>::PFQ;
  constexpr uint32_t MSG_BYTES = Q * TTB_TILE * 8;
  PipeT pipe(smem_raw, 2 * Q, 2 * PFQ, TIPS ? p.tu_stride : 0);
  const int g_ = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int k0 = group_ptr[g_], k1 = group_ptr[g_ + 1];
  const int n_chunks = k1 - k0;
  const long long a0 = (long long)tile * TTB_TILE;
  const int cols = (int)min((long long)TTB_TILE, p.ld - a0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  if (tid == 0) pipe.init();
  __syncthreads();
  unsigned long long* tr = trace_begin(p, pipe.done + 7, 1, n_chunks);
  trace_ev(tr, 0);

  SelfFed<PipeT> feed;
  Chunk nextc = load_chunk_global(chunks + k0);
  auto issue_next = [&]() {   // warp 0
    const int u = feed.issued;
    if (u >= n_chunks) return;
    const Chunk c = nextc;
    nextc = load_chunk_global(chunks + k0 + min(u + 1, n_chunks - 1));
    const int s = feed.pc.s;
    pipe.producer_acquire(feed.pc, u);
    uint64_t* bar = pipe.full + s;
    const int src = c.src0;
    const bool first = c.flags & 1;
    const bool skip = !TIPS && (p.dbg & 2);
    if (lane == 0) {
      uint32_t bytes = 32u + (first ? MSG_BYTES : 0u) + 2 * PFQ * 8u;
      bytes += (src >= 0) ? (uint32_t)(MSG_BYTES + cols) : (uint32_t)(2 * cols + p.tu_stride * 8);
      mbar_arrive_expect_tx(bar, skip ? 32u : bytes);
      tma_load_1d((void*)pipe.desc(s), chunks + k0 + u, 32, bar);
    }
    __syncwarp();
    if (!skip) {
      if (first && lane == 1) tma_load_1d(pipe.rows(s), p.M + msg_off<Q>(p, c.out, a0), MSG_BYTES, bar);
      if (lane == 2) tma_load_1d(pipe.P(s), p.Pf + (size_t)c.cnode0 * TTB_PF_STRIDE, 2 * PFQ * 8, bar);
      if (src >= 0) {
        if (lane == 3) tma_load_1d(pipe.rows(s) + Q * TTB_TILE, p.S + msg_off<Q>(p, src, a0), MSG_BYTES, bar);
        if (lane == 4) tma_load_1d(pipe.oidx(s), p.idx + (size_t)src * p.ld + a0, cols, bar);
      } else if (TIPS) {
        const int row = -1 - src;
        if (lane == 3) tma_load_1d(pipe.codes(s), p.codes + (size_t)row * p.ld + a0, cols, bar);
        if (lane == 4) tma_load_1d(pipe.oidx(s), p.idxtip + (size_t)row * p.ld + a0, cols, bar);
        if (lane == 5) tma_load_1d(pipe.TU(s), p.TU + (size_t)row * p.tu_stride, p.tu_stride * 8, bar);
      }
    }
    feed.pc.advance();
    ++feed.issued;
  };
  pdl_wait();
  trace_ev(tr, 1);
  if (warp == 0)
    for (int i = 0; i < Cfg::STAGES - 1; ++i) issue_next();

  typename PipeT::Cursor cur;
  const int g = lane >> 2, c4 = lane & 3;
  const int pat0 = warp * (MT * 8) + g;
  double Mp[MT][KS];
  unsigned int ndiff = 0, ndiff_tip = 0;
  // Software pipeline over chunks: the normalisation / argmax / stores of chunk u - 1 ("epilogue") are issued together with
  // the first product of chunk u, so a warp's stream of DMMAs is not interrupted by them (block timelines showed the
  // tensor pipe idle during a lock-step epilogue of all warps: 0.6 of 2.9 us per chunk).  Carried: the unnormalised
  // profile pr = S_c * msg, the previous states and the child's slot.
  double pr[MT][KS];
  uint8_t old[MT];
  int psrc = 0;
  bool pvalid = false;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    old[mt] = 0;
#pragma unroll
    for (int k = 0; k < KS; ++k) pr[mt][k] = 1.0;
  }
  auto epilogue = [&]() {   // of the chunk whose results are carried in pr / old / psrc
    const bool ptip = TIPS && psrc < 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const double z = quad_sum(slot_sum<KS>(pr[mt]));
      // first maximum of the unnormalised profile (the normaliser is positive), in parallel with the normaliser's chain
      // (the entries are >= 0, so their order is the order of their bit patterns as integers: the compares run on the
      // integer pipe instead of queueing behind the DMMAs on the fp64 pipe)
      long long bv = -1;
      int best = 0;
#pragma unroll
      for (int k = 0; k < KS; ++k) {   // slots are in increasing state order
        const int i = mma_state(k, c4);
        const long long v = __double_as_longlong(pr[mt][k]);
        const bool take = (i < Q) & (v > bv);
        bv = take ? v : bv;
        best = take ? i : best;
      }
#pragma unroll
      for (int d = 1; d <= 2; d <<= 1) {       // over the four lanes of the pattern
        const long long ov = __shfl_xor_sync(0xffffffffu, bv, d);
        const int ob = __shfl_xor_sync(0xffffffffu, best, d);
        const bool take = (ov > bv) | ((ov == bv) & (ob < best));
        bv = take ? ov : bv;
        best = take ? ob : best;
      }
      const double inv = fast_rcp(z);
      const long long a = a0 + pat0 + mt * 8;
      const bool act = pvalid && a < p.Lp;
      double* __restrict__ out;
      uint8_t* ip;
      if (ptip) {
        const int row = -1 - psrc;
        out = p.Mtip + msg_off<Q>(p, row, act ? a : a0);
        ip = p.idxtip + (size_t)row * p.ld + a;
      } else {
        out = p.M + msg_off<Q>(p, psrc, act ? a : a0);
        ip = p.idx + (size_t)psrc * p.ld + a;
      }
      if (act) {
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int i = mma_state(k, c4);
          if (i < Q) out[i * TTB_TILE] = pr[mt][k] * inv;
        }
      }
      if (act && c4 == 0) {
        if (count_diff) {
          const unsigned int ch = (old[mt] != (uint8_t)best);
          if (ptip) ndiff_tip += ch; else ndiff += ch;
        }
        *ip = (uint8_t)best;
      }
    }
  };
  for (int u = 0; u < n_chunks; ++u) {
    const int s = cur.s;
    if (warp == 0) issue_next();   // refills the stage of chunk u - 1
    pipe.consumer_wait(cur);
    trace_ev(tr, 2 + 3 * u);
    const Chunk c = load_chunk_smem(pipe.desc(s));
    const int src = c.src0;
    const bool tip = TIPS && src < 0;
    if (p.dbg & 1) {
      pipe.consumer_release(cur);
      cur.advance();
      continue;
    }
    {
      const bool first = c.flags & 1;
      const double* m = pipe.rows(s) + pat0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int j = mma_state(k, c4);
          if (first) Mp[mt][k] = (j < Q) ? at_least(m[j * TTB_TILE + mt * 8], TTB_TINY) : 0.0;
        }
    }
    double Sc[MT][KS], U[MT][KS];
    if (tip) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int code = (a0 + pat0 + mt * 8 < p.Lp) ? pipe.codes(s)[pat0 + mt * 8] : 0;
        const double* tu = pipe.TU(s) + code * Q;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int j = mma_state(k, c4);
          U[mt][k] = (j < Q) ? tu[j] : 1.0;
          Sc[mt][k] = (j < Q) ? __ldg(p.code_prof + code * Q + j) : 0.0;
        }
      }
      epilogue();
    } else {
      const double* rows = pipe.rows(s) + Q * TTB_TILE + pat0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int i = mma_state(k, c4);
          Sc[mt][k] = (i < Q) ? rows[i * TTB_TILE + mt * 8] : 0.0;
        }
      mma_product<MT, KS>(Sc, reinterpret_cast<const double2*>(pipe.P(s)) + lane, U);
      epilogue();   // of chunk u - 1: independent of the product above, the scheduler interleaves the two
    }
    // outside message O ~ max(TINY, profile_parent) / U (treeanc.py:895-899).  It is NOT normalised here: the product below
    // is linear in O and the profile is normalised at the end, so the factor cancels (U >= S_max * min P keeps O far from
    // the overflow range; a vanished U gives inf/NaN in either form).  Padded states stay 0.
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int j = mma_state(k, c4);
        U[mt][k] = (j < Q) ? Mp[mt][k] * fast_rcp1(U[mt][k]) : 0.0;   // 1e-12 relative: far inside the profiles' 1e-6
      }
    mma_product<MT, KS>(U, reinterpret_cast<const double2*>(pipe.P(s) + PFQ) + lane, pr);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) old[mt] = pipe.oidx(s)[pat0 + mt * 8];
    pipe.consumer_release(cur);   // all shared-memory reads of this stage are done
    trace_ev(tr, 3 + 3 * u);
    cur.advance();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int k = 0; k < KS; ++k) pr[mt][k] *= Sc[mt][k];
    psrc = src;
    pvalid = true;
    trace_ev(tr, 4 + 3 * u);
  }
  epilogue();
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
// Postorder level 1 (every child is a tip), see post_leaf_level_kernel: a pure write stream of q doubles per
// (node, pattern) behind a table lookup per tip.  Same lane layout as the kernels above (four lanes per pattern, KS
// states each): four times the threads of the one-thread-per-pattern kernel for the same bytes, 64-byte store segments.
// Block = (run of nodes, 128-pattern tile), 16 warps x 8 patterns.
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(512) post_leaf_mma_kernel(TtbDev p, const TtbChunk* __restrict__ chunks, const int* __restrict__ group_ptr,
                                                            int tiles, int fbase) {
  constexpr int KS = MmaQ<Q>::KS;
  const int g_ = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c4 = lane & 3;
  const long long a = (long long)tile * TTB_TILE + warp * 8 + g;
  pdl_launch_dependents();
  const int k0 = group_ptr[g_], k1 = group_ptr[g_ + 1];
  const bool act = a < p.Lp;
  const long long al = act ? a : 0;   // past the alignment: compute on pattern 0, store nothing
  pdl_wait();   // the tip tables come from the preceding kernel
  double X[KS];
  double Facc = 0.0, Zprod = 1.0;
  int scale = 0, seen = 0;
  Chunk c = load_chunk_global(chunks + k0);
  int code = __ldg(p.codes + (size_t)(-1 - c.src0) * p.ld + al);
  for (int k = k0; k < k1; ++k) {
    Chunk cn = c;
    int ncode = 0;
    if (k + 1 < k1) {   // descriptor and code byte of the next chunk are requested before this one is processed
      cn = load_chunk_global(chunks + k + 1);
      ncode = __ldg(p.codes + (size_t)(-1 - cn.src0) * p.ld + al);
    }
    if (c.flags & 1) {
#pragma unroll
      for (int s = 0; s < KS; ++s) X[s] = 1.0;
      scale = 0;
      seen = 0;
    }
    const double* tu = p.TU + (size_t)(-1 - c.src0) * p.tu_stride + code * Q;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int j = mma_state(s, c4);
      X[s] *= (j < Q) ? __ldg(tu + j) : 0.0;
    }
    if (++seen > 2) {
      double mx = X[0];
#pragma unroll
      for (int s = 1; s < KS; ++s) mx = fmax(mx, X[s]);
      mx = quad_max(mx);
      if (mx < 0x1p-256 && mx > 0.0) {
#pragma unroll
        for (int s = 0; s < KS; ++s) X[s] *= 0x1p+256;
        ++scale;
      }
    }
    if (c.flags & 2) {
      const double Z = quad_sum(slot_sum<KS>(X));
      const double inv = fast_rcp(Z);
      if (act) {
        double* __restrict__ so = p.S + msg_off<Q>(p, c.out, a);
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const int j = mma_state(s, c4);
          if (j < Q) so[j * TTB_TILE] = X[s] * inv;
        }
      }
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
    code = ncode;
  }
  if (act && c4 == 0) p.Fpart[(size_t)(fbase + g_) * p.ld + a] = Facc + log(Zprod);
}
