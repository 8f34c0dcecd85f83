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
// The B fragments of every branch are laid out in this order by pfrag_kernel (36 fragments x 32 lanes, zero-padded) so a
// stage receives them with one bulk copy and a lane reads its element conflict-free.
//
// Reference semantics are those of post_level_kernel / pre_level_kernel (treeanc.py:857-930); summation order differs
// (tolerances: log-LH 1e-9 relative, profiles 1e-6).  Single model, double storage, no masks, no joint pass: those keep
// the one-thread-per-pattern kernels.
#pragma once
#include "ttb_kernels.cuh"

#define TTB_MMA_NT 3                         // n-tiles of 8 states
#define TTB_MMA_KS 6                         // k-steps of 4 states (= register slots per lane and pattern)
#define TTB_MMA_NF (TTB_MMA_KS * TTB_MMA_NT) // fragments per product
static_assert(TTB_PF_STRIDE == 2 * TTB_MMA_NF * 32, "fragment-ordered exp(Qt): two products");

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// 1/a without the IEEE slow path: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps -> within 1 ulp for normal a;
// 0 -> NaN like the exact quotient's inf * 0 further down, so a vanished message surfaces the same way.
__device__ __forceinline__ double fast_rcp(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
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
__device__ __forceinline__ int mma_state(int k, int c) { return 8 * (k >> 1) + 2 * c + (k & 1); }

// exp(Qt) of every branch in fragment order (see the header comment); thread = (node, fragment, lane).
template <int Q>
__global__ void pfrag_kernel(TtbDev p, double* __restrict__ Pf) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.n_nodes * TTB_PF_STRIDE) return;
  const int node = (int)(gid / TTB_PF_STRIDE), rem = (int)(gid % TTB_PF_STRIDE);
  // fragments are stored in pairs (one 16-byte load per lane for two fragments): rem = ((f/2)*32 + lane)*2 + f%2
  const int f = (rem >> 6) * 2 + (rem & 1), lane = (rem >> 1) & 31, g = lane >> 2, c = lane & 3;
  const int which = f / TTB_MMA_NF, ff = f % TTB_MMA_NF, kap = ff / TTB_MMA_NT, ntp = ff % TTB_MMA_NT;
  const int kidx = mma_state(kap, c), nidx = 8 * ntp + g;
  const int i = which ? nidx : kidx, j = which ? kidx : nidx;
  Pf[gid] = (i < Q && j < Q) ? p.P[(size_t)node * p.pq + i * Q + j] : 0.0;
}

// C[mt][.] = A[mt][.] x B for the MT m-tiles of a warp; frag = this lane's column of the 9 fragment pairs of one product.
template <int MT>
__device__ __forceinline__ void mma_product(const double (&A)[MT][TTB_MMA_KS], const double2* __restrict__ frag, double (&C)[MT][TTB_MMA_KS]) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int k = 0; k < TTB_MMA_KS; ++k) C[mt][k] = 0.0;
#pragma unroll
  for (int pp = 0; pp < TTB_MMA_NF / 2; ++pp) {
    const double2 b = frag[pp * 32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int f = 2 * pp + h, kap = f / TTB_MMA_NT, ntp = f % TTB_MMA_NT;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) dmma884(C[mt][2 * ntp], C[mt][2 * ntp + 1], A[mt][kap], h ? b.y : b.x);
    }
  }
}

template <int Q, int NW>
struct MmaCfg {
  static_assert(Q > 16 && Q <= 24, "three n-tiles");
  static_assert(16 % NW == 0, "whole m-tiles per warp");
  static constexpr int MT = TTB_TILE / 8 / NW;        // m-tiles (8 patterns) per warp
  static constexpr int THREADS = NW * 32 + 32;        // + producer warp
  static constexpr int MAXREG = NW == 16 ? 120 : (NW == 8 ? 112 : 200);   // 1 / 2 / 2 resident blocks per SM
  using PipeT = Pipe<Q, 2, NW * 32>;
};

// ---------------------------------------------------------------------------------------
// Postorder level (A3-A5), see post_level_kernel.  Block = (run of nodes, 128-pattern tile); NW pattern warps + one
// producer warp; one child per chunk.
// ---------------------------------------------------------------------------------------
template <int Q, int NW>
__global__ void __launch_bounds__(MmaCfg<Q, NW>::THREADS) __maxnreg__((MmaCfg<Q, NW>::MAXREG)) post_level_mma_kernel(TtbDev p, const TtbChunk* __restrict__ chunks,
                                                                                 const int* __restrict__ group_ptr, int tiles, int fbase) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Cfg = MmaCfg<Q, NW>;
  using PipeT = typename Cfg::PipeT;
  constexpr int MT = Cfg::MT;
  constexpr uint32_t MSG_BYTES = Q * TTB_TILE * 8;
  constexpr int PFD = TTB_MMA_NF * 32;   // only the first product's fragments
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

  typename PipeT::Cursor cur;
  if (warp == NW) {   // producer warp
    Chunk c = load_chunk_global(chunks + k0);
    pdl_wait();
    for (int u = 0; u < n_chunks; ++u) {
      const Chunk cn = load_chunk_global(chunks + k0 + min(u + 1, n_chunks - 1));
      const int s = cur.s;
      pipe.producer_acquire(cur, u);
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
      cur.advance();
      c = cn;
    }
    return;
  }
  pdl_wait();
  const int g = lane >> 2, c4 = lane & 3;
  const int pat0 = warp * (MT * 8) + g;   // pattern (within the tile) of m-tile 0
  double X[MT][TTB_MMA_KS];
  double Facc[MT], Zprod[MT];
  int scale[MT];
  int seen = 0;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) { Facc[mt] = 0.0; Zprod[mt] = 1.0; scale[mt] = 0; }
  for (int u = 0; u < n_chunks; ++u) {
    const int s = cur.s;
    pipe.consumer_wait(cur);
    const Chunk c = load_chunk_smem(pipe.desc(s));
    if (c.flags & 1) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) X[mt][k] = 1.0;
        scale[mt] = 0;
      }
      seen = 0;
    }
    if (p.dbg & 1) {
      pipe.consumer_release(cur);
      cur.advance();
      continue;
    }
    double U[MT][TTB_MMA_KS];
    if (c.src0 >= 0) {
      double A[MT][TTB_MMA_KS];
      const double* rows = pipe.rows(s) + pat0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) {
          const int i = mma_state(k, c4);
          A[mt][k] = (i < Q) ? rows[i * TTB_TILE + mt * 8] : 0.0;
        }
      mma_product<MT>(A, reinterpret_cast<const double2*>(pipe.P(s)) + lane, U);
    } else {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int code = (a0 + pat0 + mt * 8 < p.Lp) ? pipe.codes(s)[pat0 + mt * 8] : 0;   // past the alignment: any valid row
        const double* tu = pipe.TU(s) + code * Q;
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) {
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
      for (int k = 0; k < TTB_MMA_KS; ++k) X[mt][k] *= U[mt][k];
      if (seen > 2) {   // polytomy: keep the running product in range (exact scaling)
        double mx = X[mt][0];
#pragma unroll
        for (int k = 1; k < TTB_MMA_KS; ++k) mx = fmax(mx, X[mt][k]);
        mx = quad_max(mx);
        if (mx < 0x1p-256 && mx > 0.0) {
#pragma unroll
          for (int k = 0; k < TTB_MMA_KS; ++k) X[mt][k] *= 0x1p+256;
          ++scale[mt];
        }
      }
    }
    if (c.flags & 2) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // padded states hold 0 after the first child
        const double Z = quad_sum(((X[mt][0] + X[mt][1]) + (X[mt][2] + X[mt][3])) + (X[mt][4] + X[mt][5]));
        const double inv = fast_rcp(Z);
        const long long a = a0 + pat0 + mt * 8;
        if (a < p.Lp) {
          double* __restrict__ so = p.S + msg_off<Q>(p, c.out, a);
#pragma unroll
          for (int k = 0; k < TTB_MMA_KS; ++k) {
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
  constexpr uint32_t MSG_BYTES = Q * TTB_TILE * 8;
  PipeT pipe(smem_raw, 2 * Q, TTB_PF_STRIDE, TIPS ? p.tu_stride : 0);
  const int g_ = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int k0 = group_ptr[g_], k1 = group_ptr[g_ + 1];
  const int n_chunks = k1 - k0;
  const long long a0 = (long long)tile * TTB_TILE;
  const int cols = (int)min((long long)TTB_TILE, p.ld - a0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  if (tid == 0) pipe.init();
  __syncthreads();

  typename PipeT::Cursor cur;
  if (warp == NW) {   // producer warp
    Chunk c = load_chunk_global(chunks + k0);
    pdl_wait();
    for (int u = 0; u < n_chunks; ++u) {
      const Chunk cn = load_chunk_global(chunks + k0 + min(u + 1, n_chunks - 1));
      const int s = cur.s;
      pipe.producer_acquire(cur, u);
      uint64_t* bar = pipe.full + s;
      const int src = c.src0;
      const bool first = c.flags & 1;
      const bool skip = !TIPS && (p.dbg & 2);
      if (lane == 0) {
        uint32_t bytes = 32u + (first ? MSG_BYTES : 0u) + TTB_PF_STRIDE * 8u;
        bytes += (src >= 0) ? (uint32_t)(MSG_BYTES + cols) : (uint32_t)(2 * cols + p.tu_stride * 8);
        mbar_arrive_expect_tx(bar, skip ? 32u : bytes);
        tma_load_1d((void*)pipe.desc(s), chunks + k0 + u, 32, bar);
      }
      __syncwarp();
      if (skip) {
        cur.advance();
        c = cn;
        continue;
      }
      if (first && lane == 1) tma_load_1d(pipe.rows(s), p.M + msg_off<Q>(p, c.out, a0), MSG_BYTES, bar);
      if (lane == 2) tma_load_1d(pipe.P(s), p.Pf + (size_t)c.cnode0 * TTB_PF_STRIDE, TTB_PF_STRIDE * 8, bar);
      if (src >= 0) {
        if (lane == 3) tma_load_1d(pipe.rows(s) + Q * TTB_TILE, p.S + msg_off<Q>(p, src, a0), MSG_BYTES, bar);
        if (lane == 4) tma_load_1d(pipe.oidx(s), p.idx + (size_t)src * p.ld + a0, cols, bar);
      } else if (TIPS) {
        const int row = -1 - src;
        if (lane == 3) tma_load_1d(pipe.codes(s), p.codes + (size_t)row * p.ld + a0, cols, bar);
        if (lane == 4) tma_load_1d(pipe.oidx(s), p.idxtip + (size_t)row * p.ld + a0, cols, bar);
        if (lane == 5) tma_load_1d(pipe.TU(s), p.TU + (size_t)row * p.tu_stride, p.tu_stride * 8, bar);
      }
      cur.advance();
      c = cn;
    }
    return;
  }
  pdl_wait();
  const int g = lane >> 2, c4 = lane & 3;
  const int pat0 = warp * (MT * 8) + g;
  double Mp[MT][TTB_MMA_KS];
  unsigned int ndiff = 0, ndiff_tip = 0;
  for (int u = 0; u < n_chunks; ++u) {
    const int s = cur.s;
    pipe.consumer_wait(cur);
    const Chunk c = load_chunk_smem(pipe.desc(s));
    const int src = c.src0;
    const bool tip = TIPS && src < 0;
    if (p.dbg & 1) {
      pipe.consumer_release(cur);
      cur.advance();
      continue;
    }
    if (c.flags & 1) {
      const double* m = pipe.rows(s) + pat0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) {
          const int j = mma_state(k, c4);
          Mp[mt][k] = (j < Q) ? at_least(m[j * TTB_TILE + mt * 8], TTB_TINY) : 0.0;
        }
    }
    double Sc[MT][TTB_MMA_KS], U[MT][TTB_MMA_KS];
    if (tip) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int code = (a0 + pat0 + mt * 8 < p.Lp) ? pipe.codes(s)[pat0 + mt * 8] : 0;
        const double* tu = pipe.TU(s) + code * Q;
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) {
          const int j = mma_state(k, c4);
          U[mt][k] = (j < Q) ? tu[j] : 1.0;
          Sc[mt][k] = (j < Q) ? __ldg(p.code_prof + code * Q + j) : 0.0;
        }
      }
    } else {
      const double* rows = pipe.rows(s) + Q * TTB_TILE + pat0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) {
          const int i = mma_state(k, c4);
          Sc[mt][k] = (i < Q) ? rows[i * TTB_TILE + mt * 8] : 0.0;
        }
      if (p.dbg & 4) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int k = 0; k < TTB_MMA_KS; ++k) U[mt][k] = Sc[mt][k] + 1.0;
      } else
      mma_product<MT>(Sc, reinterpret_cast<const double2*>(pipe.P(s)) + lane, U);
    }
    // outside message O ~ max(TINY, profile_parent) / U (treeanc.py:895-899).  It is NOT normalised here: the product below
    // is linear in O and the profile is normalised at the end, so the factor cancels (U >= S_max * min P keeps O far from
    // the overflow range; a vanished U gives inf/NaN in either form).  Padded states stay 0.
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int k = 0; k < TTB_MMA_KS; ++k) {
        const int j = mma_state(k, c4);
        U[mt][k] = (j < Q) ? Mp[mt][k] * fast_rcp(U[mt][k]) : 0.0;
      }
    double msg[MT][TTB_MMA_KS];
    if (p.dbg & 4) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) msg[mt][k] = U[mt][k];
    } else
    mma_product<MT>(U, reinterpret_cast<const double2*>(pipe.P(s) + TTB_MMA_NF * 32) + lane, msg);
    uint8_t old[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) old[mt] = pipe.oidx(s)[pat0 + mt * 8];
    pipe.consumer_release(cur);   // all shared-memory reads of this stage are done
    cur.advance();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int k = 0; k < TTB_MMA_KS; ++k) msg[mt][k] *= Sc[mt][k];
      const double z = quad_sum(((msg[mt][0] + msg[mt][1]) + (msg[mt][2] + msg[mt][3])) + (msg[mt][4] + msg[mt][5]));
      // first maximum of the unnormalised profile (the normaliser is positive), in parallel with the normaliser's chain
      double bv = -1.0;
      int best = 0;
#pragma unroll
      for (int k = 0; k < TTB_MMA_KS; ++k) {   // slots are in increasing state order
        const int i = mma_state(k, c4);
        if (i < Q && msg[mt][k] > bv) { bv = msg[mt][k]; best = i; }
      }
#pragma unroll
      for (int d = 1; d <= 2; d <<= 1) {       // over the four lanes of the pattern
        const double ov = __shfl_xor_sync(0xffffffffu, bv, d);
        const int ob = __shfl_xor_sync(0xffffffffu, best, d);
        if (ov > bv || (ov == bv && ob < best)) { bv = ov; best = ob; }
      }
      const double inv = fast_rcp(z);
      const long long a = a0 + pat0 + mt * 8;
      const bool act = a < p.Lp;
      double* __restrict__ out;
      uint8_t* ip;
      if (tip) {
        const int row = -1 - src;
        out = p.Mtip + msg_off<Q>(p, row, act ? a : a0);
        ip = p.idxtip + (size_t)row * p.ld + a;
      } else {
        out = p.M + msg_off<Q>(p, src, act ? a : a0);
        ip = p.idx + (size_t)src * p.ld + a;
      }
      if (act && !(p.dbg & 8)) {
#pragma unroll
        for (int k = 0; k < TTB_MMA_KS; ++k) {
          const int i = mma_state(k, c4);
          if (i < Q) out[i * TTB_TILE] = msg[mt][k] * inv;
        }
        if (c4 == 0) {
          if (count_diff) {
            const unsigned int ch = (old[mt] != (uint8_t)best);
            if (tip) ndiff_tip += ch; else ndiff += ch;
          }
          *ip = (uint8_t)best;
        }
      }
    }
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
