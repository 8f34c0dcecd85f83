// ttb_api.cu -- host side of libttb.so: the C-ABI of include/ttb.h, device memory,
// level schedules and the CUDA-graph that covers one marginal reconstruction.
#include "../../include/ttb.h"
#include "ttb_qops.h"
#include "ttb_mma.cuh"   // TTB_PF_STRIDE, trace layout
#include "ttb_brent.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return fail(e_ == cudaErrorMemoryAllocation ? TTB_ENOMEM : TTB_ECUDA, buf_);            \
    }                                                                                         \
  } while (0)

#define TTB_FOR_EACH_Q(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(20) X(21) X(22)
#define TTB_DECL(qv) extern const TtbQOps ttb_qops_##qv;
}  // namespace
TTB_FOR_EACH_Q(TTB_DECL)
const TtbQOps* ttb_qops(int q) {
  switch (q) {
#define TTB_CASE(qv) case qv: return &ttb_qops_##qv;
    TTB_FOR_EACH_Q(TTB_CASE)
    default: return nullptr;
  }
}
namespace {

bool q_supported(int q) { return ttb_qops(q) != nullptr; }

template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }   // `delete handle` frees every device buffer, listed in ttb_destroy or not
  int alloc(size_t count) {
    if (count == n && p) return 0;
    release();
    if (count == 0) return 0;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();
      char buf[256];
      snprintf(buf, sizeof buf, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
      return fail(TTB_ENOMEM, buf);
    }
    n = count;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// Level-ordered schedule of one traversal: chunks grouped by node, nodes grouped by level.
struct Sched {
  std::vector<TtbChunk> chunks;       // all levels, node-major
  std::vector<int> node_chunk;        // first chunk of every scheduled node (+ sentinel), level-major
  std::vector<int> level_node_begin;  // offsets into node_chunk per level (+ sentinel)
  std::vector<int> group_ptr;         // built once the number of tiles is known
  std::vector<TtbLevelLaunch> launches;
  std::vector<int> dep;               // per chunk: the chunk whose output it reads (latest one), -1 = tips / the root only
  DBuf<TtbChunk> d_chunks;
  DBuf<int> d_group_ptr, d_node_chunk, d_dep;
  void clear() {
    chunks.clear(); node_chunk.clear(); level_node_begin.clear(); group_ptr.clear(); launches.clear(); dep.clear();
  }
};

}  // namespace

struct ttb_engine {
  int device = 0;
  int n_sm = 148;   // multiprocessors of this device (grid sizing)
  int q = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  // tree (host)
  int n_nodes = 0, n_int = 0, n_tips = 0;
  std::vector<int> parent, child_ptr, child_idx, tip_row, int_slot;
  std::vector<int> tip_nodes;
  // joint back-trace: non-root nodes by depth, internal only / all
  std::vector<int> jpre_int_nodes, jpre_all_nodes;
  std::vector<TtbLevelLaunch> jpre_int_levels, jpre_all_levels;
  DBuf<int> d_jpre_int, d_jpre_all;
  bool have_joint = false, have_joint_tips = false;
  Sched post, pre_int, pre_all;
  int sched_tiles = -1;  // tiles the group pointers were built for
  int n_fgroups = 0;     // postorder block runs (rows of Fpart)
  int sched_ss = 0;      // 0 single model, 1 site-specific, 2 site-specific symmetric (3 blocks/SM)
  bool prepared = false;
  // device tree
  DBuf<int> d_parent, d_child_ptr, d_child_idx, d_tip_row, d_int_slot, d_tip_nodes;
  // alignment
  long long Lp = 0, ld = 0;
  int n_codes = 0;
  DBuf<uint8_t> d_codes;
  DBuf<double> d_code_prof, d_mult;
  std::vector<double> h_mult;
  // model
  bool have_gtr = false, have_t = false;
  bool f32 = false;   // S / M / Mtip stored as float (ttb_set_message_storage)
  double mu = 1.0;
  int gap_index = -1;
  DBuf<double> d_t, d_eig, d_v, d_vinv, d_Pi, d_mu;
  // site-specific model
  bool site_specific = false;
  DBuf<double> d_ss_eig, d_ss_mu, d_ss_V, d_ss_Vinv, d_ss_Pi, d_ss_w, d_ss_grid, d_ss_E, d_ss_c, d_ss_Ec;
  bool ss_sym = false;
  DBuf<int> d_ss_lo;
  DBuf<double2> d_ss_rec;
  std::vector<double> ss_grid, h_t;
  double ss_tmax = 0.0;
  bool ss_interp_dirty = true;
  // state
  // device-side lock-step Brent (ttb_brent_*)
  DBuf<double> d_brent;        // 18 state vectors + trial lengths + objective values, n_brent entries each
  DBuf<int> d_brent_i;         // nit, nfev, flags
  DBuf<uint8_t> d_brent_active;
  DBuf<unsigned long long> d_trace;   // measurement only (TTB_TRACE)
  int n_brent = 0, brent_stage = 0, brent_maxiter = 500, brent_nb = 1;
  double brent_tol = 0.0;
  int* h_brent_flags = nullptr;   // pinned {n_active, bracket error}
  DBuf<double> d_leaf_pairs;   // cherry tables of postorder level 1 (leaf_pair_table_kernel)
  DBuf<double> d_LP, d_TL, d_Fred, d_TU, d_P, d_Pf, d_S, d_F, d_M, d_Mtip, d_LH, d_lh_partial, d_results, d_stage, d_partial;
  DBuf<uint8_t> d_TC, d_Cx, d_idx, d_idxtip, d_bstage, d_mut_state, d_aln, d_colstat, d_lut, d_constl;
  DBuf<long long> d_firstpos;
  DBuf<int> d_seqrow, d_flag;
  long long aln_rows = 0, aln_L = 0;
  DBuf<int> d_mut_node, d_mut_pos, d_ent_row, d_ent_pos;
  DBuf<long long> d_mut_offsets;
  DBuf<unsigned long long> d_mut_count;   // d_bstage: packed byte staging for contiguous H2D / D2H
  DBuf<unsigned long long> d_nd;
  DBuf<int> d_enodes, d_ekinds, d_pair_first;
  DBuf<double> d_pair_counts, d_sg_uniforms;
  DBuf<uint8_t> d_sg_states;
  // sample_from_profile=True: states of the previous pass (snapshot taken by TTB_KEEP_PREV_STATES)
  DBuf<uint8_t> d_idx_prev, d_idxtip_prev;
  DBuf<unsigned long long> d_scount;
  // per-branch masks (ARG mode)
  DBuf<int> d_mask_id;
  DBuf<uint8_t> d_masks;
  bool have_masks = false;
  bool have_prev = false, have_prev_tips = false;
  DBuf<double> d_ets, d_eout;
  double* h_results = nullptr;  // pinned {total, ndiff}
  // page-locked scratch through which small pageable inputs (branch lengths, model, multiplicities)
  // are staged, so that those uploads never block the host behind queued GPU work
  unsigned char* h_scratch = nullptr;
  size_t scratch_cap = 0, scratch_used = 0;
  cudaEvent_t scratch_ev = nullptr;
  bool have_pass = false;       // a full (non LH-only) pass has completed
  bool have_tip_pass = false;
  bool first_full = true;       // no previous state indices to diff against
  std::map<int, cudaGraphExec_t> graphs;
  std::map<int, int> graph_kernels;
  long long launches = 0;

  int tiles() const { return (int)((Lp + TTB_BLOCK - 1) / TTB_BLOCK); }
  // chunks of postorder level 1 (they come first in the schedule)
  int n_leaf_chunks() const {
    return post.level_node_begin.size() > 1 ? post.node_chunk[post.level_node_begin[1]] : 0;
  }
  // doubles to allocate for a message array of n elements in the current storage type
  size_t msg_doubles(size_t n) const { return f32 ? (n + 1) / 2 : n; }

  void drop_graphs() {
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    graphs.clear();
    graph_kernels.clear();
  }

  TtbDev dev() const {
    TtbDev d;
    d.q = q;
    d.Lp = Lp;
    d.ld = ld;
    d.tiles = tiles();
    d.n_nodes = n_nodes;
    d.n_int = n_int;
    d.n_tips = n_tips;
    d.n_codes = n_codes;
    d.gap_index = gap_index;
    d.parent = d_parent.p;
    d.child_ptr = d_child_ptr.p;
    d.child_idx = d_child_idx.p;
    d.tip_row = d_tip_row.p;
    d.int_slot = d_int_slot.p;
    d.codes = d_codes.p;
    d.code_prof = d_code_prof.p;
    d.mult = d_mult.p;
    d.t = d_t.p;
    d.eig = d_eig.p;
    d.v = d_v.p;
    d.vinv = d_vinv.p;
    d.Pi = d_Pi.p;
    d.mu = d_mu.p;
    d.site_specific = site_specific ? 1 : 0;
    d.ss_eig = d_ss_eig.p; d.ss_mu = d_ss_mu.p; d.ss_V = d_ss_V.p; d.ss_Vinv = d_ss_Vinv.p; d.ss_Pi = d_ss_Pi.p;
    d.ss_rec = d_ss_rec.p;
    d.ss_lo = d_ss_lo.p; d.ss_w = d_ss_w.p; d.ss_grid = d_ss_grid.p; d.ss_E = d_ss_E.p;
    d.ss_sym = ss_sym ? 1 : 0; d.ss_c = d_ss_c.p; d.ss_Ec = d_ss_Ec.p;
    d.mask_id = have_masks ? d_mask_id.p : nullptr;
    d.masks = have_masks ? d_masks.p : nullptr;
    d.ss_ngrid = (int)ss_grid.size();
    d.ss_tmax = ss_tmax;
    d.pq = (q * q + 1) / 2 * 2;
    d.tu_stride = (n_codes * q + 1) / 2 * 2;
    d.TU = d_TU.p;
    d.P = d_P.p;
    d.Pf = d_Pf.p;
    d.f32 = f32 ? 1 : 0;
    d.S = d_S.p;
    d.Fpart = d_F.p;
    d.Fred = d_Fred.p;
    d.n_fgroups = n_fgroups;
    d.M = d_M.p;
    d.Mtip = d_Mtip.p;
    d.idx = d_idx.p;
    d.idxtip = d_idxtip.p;
    d.LP = d_LP.p; d.TL = d_TL.p; d.TC = d_TC.p; d.Cx = d_Cx.p;
    d.LH = d_LH.p;
    d.lh_partial = d_lh_partial.p;
    d.nd_slots = d_nd.p;
    d.results = d_results.p;
    d.trace = d_trace.p;
    { static const int dbg = getenv("TTB_DBG") ? atoi(getenv("TTB_DBG")) : 0; d.dbg = dbg; }
    return d;
  }
};

namespace {

int use_device(ttb_handle h) {
  if (!h) return fail(TTB_EINVAL, "null handle");
  CK(cudaSetDevice(h->device));
  return 0;
}

template <typename T>
int upload(DBuf<T>& b, const T* src, size_t n, cudaStream_t s) {
  int rc = b.alloc(n);
  if (rc) return rc;
  if (n) CK(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
  return 0;
}

// Asynchronous H2D of a small pageable host array: copy it into the handle's page-locked scratch
// first.  The scratch is a bump allocator guarded by an event (recycled only after the copies that
// read it have executed).
int upload_small(ttb_handle h, void* dst, const void* src, size_t bytes) {
  if (!bytes) return 0;
  const size_t need = (bytes + 255) / 256 * 256;
  if (need > h->scratch_cap / 2) {   // too big for the scratch: plain (possibly blocking) copy
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
  }
  if (h->scratch_used + need > h->scratch_cap) {
    CK(cudaEventSynchronize(h->scratch_ev));
    h->scratch_used = 0;
  }
  unsigned char* p = h->h_scratch + h->scratch_used;
  memcpy(p, src, bytes);
  h->scratch_used += need;
  CK(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(h->scratch_ev, h->stream));
  return 0;
}

// Build the chunk schedule of one traversal.  `key[n]` = level of node n (height for the
// postorder, depth for the preorder), `member[n]` = node n is scheduled, `take(c)` = child c
// takes part.  Levels are emitted in increasing key order for the preorder and increasing
// height for the postorder (both = dependency order); nodes inside a level in id order.
template <typename Take>
void build_sched(ttb_handle h, const std::vector<int>& key, const std::vector<char>& member, bool out_is_node_slot,
                 Take take, Sched& sc) {
  sc.clear();
  const int n_nodes = h->n_nodes;
  int maxk = -1;
  for (int n = 0; n < n_nodes; ++n)
    if (member[n]) maxk = std::max(maxk, key[n]);
  std::vector<std::vector<int>> by_level(maxk + 1);
  for (int n = 0; n < n_nodes; ++n)
    if (member[n]) by_level[key[n]].push_back(n);
  sc.level_node_begin.push_back(0);
  for (int k = 0; k <= maxk; ++k) {
    if (by_level[k].empty()) continue;
    for (int n : by_level[k]) {
      sc.node_chunk.push_back((int)sc.chunks.size());
      std::vector<int> kids;
      for (int e = h->child_ptr[n]; e < h->child_ptr[n + 1]; ++e)
        if (take(h->child_idx[e])) kids.push_back(h->child_idx[e]);
      const int nk = (int)kids.size();
      const int CB = h->q <= 8 ? TTB_CB : 1;   // same rule as Pipe<Q>::CB
      for (int c0 = 0; c0 < nk; c0 += CB) {
        TtbChunk ch;
        memset(&ch, 0, sizeof ch);
        ch.out = h->int_slot[n];
        const int nb = std::min(CB, nk - c0);
        ch.flags = (c0 == 0 ? 1 : 0) | (c0 + CB >= nk ? 2 : 0) | (nb << 8);
        for (int b = 0; b < nb; ++b) {
          const int c = kids[c0 + b];
          ch.src[b] = h->int_slot[c] >= 0 ? h->int_slot[c] : -1 - h->tip_row[c];
          ch.cnode[b] = c;
        }
        sc.chunks.push_back(ch);
      }
    }
    sc.level_node_begin.push_back((int)sc.node_chunk.size());
  }
  sc.node_chunk.push_back((int)sc.chunks.size());
  // Producer -> consumer links between chunks (used by launches that merge several levels, see build_groups):
  // postorder (out_is_node_slot): a chunk reads the subtree profiles its internal children got in their LAST chunk;
  // preorder: a parent's first chunk reads the profile the parent received as a child of an earlier chunk.
  sc.dep.assign(sc.chunks.size(), -1);
  std::vector<int> wrote(n_nodes, -1);   // node -> chunk that wrote its message in this schedule
  if (out_is_node_slot) {
    for (size_t i = 0; i < sc.chunks.size(); ++i) {
      const TtbChunk& ch = sc.chunks[i];
      for (int b = 0; b < (ch.flags >> 8); ++b)
        if (ch.src[b] >= 0) sc.dep[i] = std::max(sc.dep[i], wrote[ch.cnode[b]]);
      if (ch.flags & 2) wrote[h->parent[ch.cnode[0]]] = (int)i;
    }
  } else {
    for (size_t i = 0; i < sc.chunks.size(); ++i) {
      const TtbChunk& ch = sc.chunks[i];
      sc.dep[i] = wrote[h->parent[ch.cnode[0]]];
      for (int b = 0; b < (ch.flags >> 8); ++b) wrote[ch.cnode[b]] = (int)i;
    }
  }
}

// Split every level into groups of consecutive nodes so that a launch has enough blocks to
// fill the GPU but every block still pipelines over several chunks.
void build_groups(Sched& sc, int tiles, bool site_specific, int ss_blocks_per_sm, bool post_order, int n_sm, bool allow_merge, bool mma) {
  sc.group_ptr.clear();
  sc.launches.clear();
  // Site-specific kernels load a per-pattern eigen-system per block (longer runs amortise it) and are
  // issue-bound, so a partly filled last wave costs its full duration: size the grid of a level to fill
  // whole waves of the 2 blocks/SM these kernels run at (3 for the symmetric variant).
  long long slots = (long long)n_sm * ss_blocks_per_sm;
  if (const char* e = getenv("TTB_SS_SLOTS")) slots = std::max(1LL, atoll(e));
  // single-model kernels: ~2 waves of 3 blocks/SM per level -- longer runs amortise a block's prologue and first
  // bulk copy (measured: 888 vs 4736 blocks per level: cfg2 0.95 -> 0.75 ms, cfg3 14.26 -> 14.13 ms, cfg4 2.33 -> 2.20 ms)
  long long target_blocks = site_specific ? slots * 4 : (long long)n_sm * 6;
  long long max_group = site_specific ? 128 : 32;
  // tensor-pipe kernels of the large alphabets (ttb_mma.cuh): ONE wave of resident blocks per level -- two blocks per SM in
  // the postorder, one in the preorder -- so that a block's start-up (~2-4 us: barriers, descriptors, first bulk copy) is
  // paid once per level (cfg4 sweep, profiles/R2p_mma_sweep.txt: 888 -> 296 / 148 blocks: 1.76 -> 1.63 ms per pass)
  if (mma) target_blocks = post_order ? (long long)n_sm * 2 : (long long)n_sm;
  // single-model kernels of the small alphabets: ONE whole wave of the 3 resident blocks per SM as well (cfg2 sweep,
  // profiles/R2aa_cfg2_grouping.txt: 296 / 444 / 592 / 888 / 1332 blocks per level -> 0.80 / 0.71 / 0.81 / 0.73 / 0.77 ms:
  // whole waves win, one beats two); big levels exceed it anyway through the run-length cap
  const bool one_wave = mma || !site_specific;
  if (!site_specific && !mma) target_blocks = (long long)n_sm * 3;
  if (const char* e = getenv("TTB_TARGET_BLOCKS")) target_blocks = std::max(1LL, atoll(e));   // tuning knobs (measurement only)
  if (const char* e = getenv("TTB_MAX_GROUP")) max_group = std::max(1LL, atoll(e));
  long long ss_waves = 4;
  if (const char* e = getenv("TTB_SS_WAVES")) ss_waves = std::max(1LL, atoll(e));
  // Merged-level launches (opt-in, TTB_MERGE_NODES=k): a run of consecutive levels with at most k nodes each becomes ONE
  // launch with one group per pattern tile -- the block walks all those levels, its producer warp honouring the chunks'
  // dependencies (Pipe::wait_done).  Parity-tested (tests/test_gpu_parity.py::test_merged_level_launches) and measured
  // (profiles/R2b_merge_threshold.json): it is SLOWER at every threshold and configuration (cfg2 0.763 -> 0.824 ms at
  // k = 24, cfg4 2.01 -> 2.61 ms, cfg3 / cfg5 unchanged): with programmatic dependent launches inside the CUDA graph a
  // small level costs ~2.2 us, while one block walking the nodes of a level one after the other needs ~0.66 us per node
  // (write -> proxy fence -> bulk read of the same tile is a latency chain).  So the default stays one launch per level.
  // Postorder level 0 (all children are tips) has its own kernel and is never merged.
  long long merge_nodes = 0;
  if (const char* e = getenv("TTB_MERGE_NODES")) merge_nodes = std::max(0LL, atoll(e));
  if (!allow_merge) merge_nodes = 0;     // the DEP kernels exist for single models, double storage, no masks
  const size_t n_levels = sc.level_node_begin.size() - 1;
  for (size_t l = 0; l < n_levels; ++l) {
    const int nb = sc.level_node_begin[l];
    size_t l1 = l + 1;
    if (merge_nodes > 0 && (l > 0 || !post_order)) {
      size_t m = l;
      while (m < n_levels && sc.level_node_begin[m + 1] - sc.level_node_begin[m] <= merge_nodes) ++m;
      if (m - l >= 2) l1 = m;
    }
    if (l1 > l + 1) {
      TtbLevelLaunch L;
      L.group_off = (int)sc.group_ptr.size();
      L.n_groups = 1;
      L.dep = 1;
      sc.group_ptr.push_back(sc.node_chunk[nb]);
      sc.group_ptr.push_back(sc.node_chunk[sc.level_node_begin[l1]]);
      sc.launches.push_back(L);
      l = l1 - 1;
      continue;
    }
    const int ne = sc.level_node_begin[l + 1];
    const int n = ne - nb;
    long long G;
    if (site_specific) {
      long long waves = ss_waves;
      for (;;) {
        const long long groups = std::max(1LL, std::min((long long)n, waves * slots / tiles));
        G = (n + groups - 1) / groups;
        if (G <= max_group) break;
        ++waves;
      }
    } else {
      G = ((long long)n * tiles + target_blocks - 1) / target_blocks;
      if (one_wave && !(post_order && l == 0 && !mma)) {
        // never more blocks than one wave holds: a few blocks beyond it would double the level's duration (the leaf level
        // of the small alphabets has its own, unpipelined kernel with 10 resident blocks per SM and keeps the plain rule)
        const long long groups_max = std::max(1LL, target_blocks / tiles);
        G = (n + groups_max - 1) / groups_max;
      }
      G = std::max(1LL, std::min(max_group, G));
    }
    TtbLevelLaunch L;
    L.group_off = (int)sc.group_ptr.size();
    L.n_groups = 0;
    L.dep = 0;
    for (int i = nb; i < ne; i += (int)G) {
      sc.group_ptr.push_back(sc.node_chunk[i]);
      ++L.n_groups;
    }
    sc.group_ptr.push_back(sc.node_chunk[ne]);
    sc.launches.push_back(L);
  }
}

// Enqueue every kernel of one pass on `s`; returns the number of kernels.
int enqueue_pass(ttb_handle h, int flags, int count_diff, cudaStream_t s, int* n_kernels, cudaEvent_t* ev = nullptr,
                 int* phase_kernels = nullptr) {
  TtbPassPlan pl;
  pl.d = h->dev();
  pl.tiles = h->tiles();
  pl.lh_only = flags & TTB_LH_ONLY;
  pl.tips = flags & TTB_RECONSTRUCT_TIPS;
  pl.count_diff = count_diff;
  pl.d_tip_nodes = h->d_tip_nodes.p;
  pl.d_post_chunks = h->post.d_chunks.p;
  pl.d_post_group_ptr = h->post.d_group_ptr.p;
  pl.d_post_dep = h->post.d_dep.p;
  pl.d_leaf_pairs = h->d_leaf_pairs.p;
  pl.n_leaf_chunks = h->n_leaf_chunks();
  pl.d_post_node_chunk = h->post.d_node_chunk.p;
  pl.n_post_leaf_nodes = h->post.level_node_begin.size() > 1 ? h->post.level_node_begin[1] : 0;
  pl.post_levels = h->post.launches.data();
  pl.n_post_levels = (int)h->post.launches.size();
  const Sched& pre = pl.tips ? h->pre_all : h->pre_int;
  pl.d_pre_chunks = pre.d_chunks.p;
  pl.d_pre_dep = pre.d_dep.p;
  pl.d_pre_group_ptr = pre.d_group_ptr.p;
  pl.pre_levels = pre.launches.data();
  pl.n_pre_levels = (int)pre.launches.size();
  *n_kernels = ttb_qops(h->q)->enqueue_pass(pl, s, ev, phase_kernels);
  return 0;
}

void fill_plan(ttb_handle h, TtbPassPlan& pl, bool tips, int count_diff) {
  pl.d = h->dev();
  pl.tiles = h->tiles();
  pl.lh_only = false;
  pl.tips = tips;
  pl.count_diff = count_diff;
  pl.d_tip_nodes = h->d_tip_nodes.p;
  pl.d_post_chunks = h->post.d_chunks.p;
  pl.d_post_group_ptr = h->post.d_group_ptr.p;
  pl.d_post_dep = h->post.d_dep.p;
  pl.d_leaf_pairs = h->d_leaf_pairs.p;
  pl.n_leaf_chunks = h->n_leaf_chunks();
  pl.d_post_node_chunk = h->post.d_node_chunk.p;
  pl.n_post_leaf_nodes = h->post.level_node_begin.size() > 1 ? h->post.level_node_begin[1] : 0;
  pl.post_levels = h->post.launches.data();
  pl.n_post_levels = (int)h->post.launches.size();
  pl.d_pre_chunks = nullptr; pl.d_pre_group_ptr = nullptr; pl.d_pre_dep = nullptr; pl.pre_levels = nullptr; pl.n_pre_levels = 0;
  pl.d_jpre_nodes = tips ? h->d_jpre_all.p : h->d_jpre_int.p;
  pl.jpre_levels = tips ? h->jpre_all_levels.data() : h->jpre_int_levels.data();
  pl.n_jpre_levels = (int)(tips ? h->jpre_all_levels.size() : h->jpre_int_levels.size());
}

int ensure_state(ttb_handle h, bool tips) {
  const size_t q = h->q, ld = h->ld;
  int rc;
  const size_t pq = (q * q + 1) / 2 * 2, tus = ((size_t)h->n_codes * q + 1) / 2 * 2;
  if ((rc = h->d_P.alloc((size_t)h->n_nodes * pq))) return rc;
  {
    // large alphabets: exp(Qt) also in mma-fragment order for the tensor-pipe level kernels (ttb_mma.cuh); TTB_NO_MMA=1
    // keeps the one-thread-per-pattern kernels (A/B measurements)
    const char* mg = getenv("TTB_MERGE_NODES");   // merged-level launches exist for the one-thread-per-pattern kernels only
    const bool mma = q > 8 && !h->site_specific && !getenv("TTB_NO_MMA") && !(mg && atoll(mg) > 0);
    const double* before = h->d_Pf.p;
    if ((rc = h->d_Pf.alloc(mma ? (size_t)h->n_nodes * TTB_PF_STRIDE : 0))) return rc;
    if (before != h->d_Pf.p) h->drop_graphs();
  }
  if ((rc = h->d_TU.alloc((size_t)h->n_tips * tus))) return rc;
  // grouping mode: 0 single model, 1 site-specific, 2 site-specific symmetric; +4: no merged-level launches (masks / float storage);
  // +8: tensor-pipe level kernels (large alphabets)
  const int ss_mode = (h->site_specific ? (h->ss_sym ? 2 : 1) : 0) | ((h->have_masks || h->f32) ? 4 : 0) | ((h->d_Pf.p && !h->have_masks) ? 8 : 0);
  if (h->sched_tiles != h->tiles() || h->sched_ss != ss_mode) {
    for (Sched* sc : {&h->post, &h->pre_int, &h->pre_all}) {
      build_groups(*sc, h->tiles(), h->site_specific, (ss_mode & 3) == 2 ? 3 : 2, sc == &h->post, h->n_sm, ss_mode == 0, (ss_mode & 8) != 0);
      if ((rc = upload(sc->d_group_ptr, sc->group_ptr.data(), sc->group_ptr.size(), h->stream))) return rc;
    }
    CK(cudaStreamSynchronize(h->stream));
    h->n_fgroups = 0;
    for (const TtbLevelLaunch& L : h->post.launches) h->n_fgroups += L.n_groups;
    h->sched_tiles = h->tiles();
    h->sched_ss = ss_mode;
    h->drop_graphs();
  }
  // cherry tables (nucleotide-sized alphabets, single model): n_codes^2 entries per level-1 chunk, at most 2 GB
  {
    const size_t per = (size_t)h->n_codes * h->n_codes * TTB_PAIR_STRIDE(q);
    const size_t want = (q <= 8 && !h->site_specific && per * 8 <= 32 * 1024) ? (size_t)h->n_leaf_chunks() * per : 0;
    const bool on = want && want * 8 <= ((size_t)2 << 30) && !getenv("TTB_NO_PAIR_TABLES");
    const double* before = h->d_leaf_pairs.p;
    if ((rc = h->d_leaf_pairs.alloc(on ? want : 0))) return rc;
    if (before != h->d_leaf_pairs.p) h->drop_graphs();
  }
  if ((rc = h->d_F.alloc((size_t)h->n_fgroups * ld))) return rc;
  if ((rc = h->d_Fred.alloc((size_t)TTB_FLANES * ld))) return rc;
  if (!h->prepared) {
    const int e = ttb_qops(h->q)->prepare(h->dev());
    if (e) return fail(TTB_ECUDA, std::string("cudaFuncSetAttribute(shared memory) failed: ") + cudaGetErrorString((cudaError_t)e));
    h->prepared = true;
  }
  const size_t msg = (size_t)h->tiles() * q * TTB_TILE;   // one node's tile-blocked message: [tiles][q][128]
  if ((rc = h->d_S.alloc(h->msg_doubles((size_t)h->n_int * msg)))) return rc;
  if ((rc = h->d_LH.alloc(ld))) return rc;
  if ((rc = h->d_lh_partial.alloc(h->tiles()))) return rc;
  if ((rc = h->d_nd.alloc(1024))) return rc;
  if ((rc = h->d_results.alloc(4))) return rc;
  return 0;
}

int ensure_preorder_state(ttb_handle h, bool tips) {
  const size_t q = h->q, ld = h->ld;
  int rc;
  if (!h->d_M.p) {
    if ((rc = h->d_M.alloc(h->msg_doubles((size_t)h->n_int * h->tiles() * q * TTB_TILE)))) return rc;
    if (!h->d_idx.p) {      // a joint pass may already have left states here: they are the "previous" ones
      if ((rc = h->d_idx.alloc((size_t)h->n_int * ld))) return rc;
      CK(cudaMemsetAsync(h->d_idx.p, 0xff, h->d_idx.bytes(), h->stream));
    }
    h->drop_graphs();
  }
  if (tips && !h->d_Mtip.p) {
    if ((rc = h->d_Mtip.alloc(h->msg_doubles((size_t)h->n_tips * h->tiles() * q * TTB_TILE)))) return rc;
    if (!h->d_idxtip.p) {
      if ((rc = h->d_idxtip.alloc((size_t)h->n_tips * ld))) return rc;
      CK(cudaMemsetAsync(h->d_idxtip.p, 0xff, h->d_idxtip.bytes(), h->stream));
    }
    h->drop_graphs();
  }
  return 0;
}

// reconstructed states exist after a marginal OR a joint pass
int check_states(ttb_handle h) {
  if (!h->have_pass && !h->have_joint) return fail(TTB_EINVAL, "no reconstruction has been run yet (call ttb_marginal or ttb_joint first)");
  return 0;
}

int check_ready(ttb_handle h, bool need_pass) {
  if (!h->n_nodes) return fail(TTB_EINVAL, "ttb_set_tree has not been called");
  if (!h->Lp) return fail(TTB_EINVAL, "ttb_set_patterns has not been called");
  if (!h->have_gtr) return fail(TTB_EINVAL, "ttb_set_gtr has not been called");
  if (!h->have_t) return fail(TTB_EINVAL, "ttb_set_branch_lengths has not been called");
  if (need_pass && !h->have_pass) return fail(TTB_EINVAL, "no marginal reconstruction has been run yet (call ttb_marginal first)");
  return 0;
}

}  // namespace

extern "C" {

const char* ttb_last_error(void) { return g_err.c_str(); }
int ttb_version(void) { return 100; }
int ttb_supports_n_states(int n_states) { return q_supported(n_states) ? 1 : 0; }

int ttb_create(ttb_handle* out, int device, int n_states) {
  if (!out) return fail(TTB_EINVAL, "out is null");
  if (!q_supported(n_states)) {
    char buf[128];
    snprintf(buf, sizeof buf, "no kernels compiled for n_states=%d (supported: 2..8, 20..22)", n_states);
    return fail(TTB_EUNSUPPORTED, buf);
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(TTB_EINVAL, "no such CUDA device");
  CK(cudaSetDevice(device));
  ttb_engine* h = new ttb_engine();
  h->device = device;
  h->q = n_states;
  CK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device));
  CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  CK(cudaMallocHost(&h->h_results, 4 * sizeof(double)));
  h->scratch_cap = 16u << 20;
  CK(cudaMallocHost(&h->h_scratch, h->scratch_cap));
  CK(cudaEventCreateWithFlags(&h->scratch_ev, cudaEventDisableTiming));
  *out = h;
  return 0;
}

int ttb_destroy(ttb_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->drop_graphs();
  DBuf<int>* ib[] = {&h->d_parent, &h->d_child_ptr, &h->d_child_idx, &h->d_tip_row, &h->d_int_slot, &h->d_tip_nodes,
                     &h->d_enodes, &h->d_ekinds, &h->post.d_group_ptr, &h->pre_int.d_group_ptr, &h->pre_all.d_group_ptr, &h->post.d_node_chunk,
                     &h->post.d_dep, &h->pre_int.d_dep, &h->pre_all.d_dep};
  for (auto* b : ib) b->release();
  DBuf<double>* db[] = {&h->d_code_prof, &h->d_mult, &h->d_t, &h->d_eig, &h->d_v, &h->d_vinv, &h->d_Pi, &h->d_mu, &h->d_LP, &h->d_TL, &h->d_ss_eig, &h->d_ss_mu, &h->d_ss_V, &h->d_ss_Vinv, &h->d_ss_Pi,
                        &h->d_ss_w, &h->d_ss_grid, &h->d_ss_E, &h->d_TU, &h->d_P, &h->d_Pf, &h->d_S,
                        &h->d_F, &h->d_Fred, &h->d_M, &h->d_Mtip, &h->d_LH, &h->d_lh_partial, &h->d_results, &h->d_stage,
                        &h->d_partial, &h->d_ets, &h->d_eout};
  for (auto* b : db) b->release();
  h->d_codes.release();
  h->d_ss_lo.release();
  h->d_ss_rec.release();
  h->d_pair_first.release();
  h->d_pair_counts.release();
  h->d_sg_states.release();
  h->d_sg_uniforms.release();
  h->d_idx_prev.release(); h->d_idxtip_prev.release(); h->d_scount.release();
  h->d_ss_c.release(); h->d_ss_Ec.release();
  h->d_mask_id.release(); h->d_masks.release();
  h->post.d_chunks.release(); h->pre_int.d_chunks.release(); h->pre_all.d_chunks.release();
  h->d_idx.release();
  h->d_idxtip.release();
  h->d_bstage.release();
  h->d_TC.release(); h->d_Cx.release(); h->d_jpre_int.release(); h->d_jpre_all.release();
  h->d_mut_state.release(); h->d_mut_node.release(); h->d_mut_pos.release(); h->d_ent_row.release(); h->d_ent_pos.release();
  h->d_mut_count.release(); h->d_mut_offsets.release();
  h->d_aln.release(); h->d_colstat.release(); h->d_lut.release(); h->d_constl.release(); h->d_firstpos.release();
  h->d_seqrow.release(); h->d_flag.release();
  h->d_nd.release();
  if (h->h_brent_flags) cudaFreeHost(h->h_brent_flags);
  if (h->h_results) cudaFreeHost(h->h_results);
  if (h->h_scratch) cudaFreeHost(h->h_scratch);
  if (h->scratch_ev) cudaEventDestroy(h->scratch_ev);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return 0;
}

int ttb_set_stream(ttb_handle h, void* cuda_stream) {
  if (int rc = use_device(h)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}

int ttb_set_message_storage(ttb_handle h, int32_t storage) {
  if (int rc = use_device(h)) return rc;
  if (storage != TTB_STORAGE_F64 && storage != TTB_STORAGE_F32) return fail(TTB_EINVAL, "ttb_set_message_storage: unknown storage type");
  const bool f32 = storage == TTB_STORAGE_F32;
  if (f32 && h->q > 8) return fail(TTB_EUNSUPPORTED, "float message storage is compiled for alphabets of up to 8 states");
  if (f32 == h->f32) return 0;
  CK(cudaStreamSynchronize(h->stream));
  h->f32 = f32;
  h->d_S.release(); h->d_M.release(); h->d_Mtip.release();   // reallocated in the new type by the next pass
  h->have_pass = h->have_tip_pass = false;
  h->have_joint = h->have_joint_tips = false;
  h->drop_graphs();
  return 0;
}

int ttb_set_tree(ttb_handle h, int32_t n_nodes, const int32_t* parent, const int32_t* child_ptr,
                 const int32_t* child_idx, const int32_t* tip_row) {
  if (int rc = use_device(h)) return rc;
  if (n_nodes < 2 || !parent || !child_ptr || !child_idx || !tip_row) return fail(TTB_EINVAL, "ttb_set_tree: bad arguments");
  if (parent[0] != -1) return fail(TTB_EINVAL, "ttb_set_tree: node 0 must be the root (parent -1)");
  if (child_ptr[0] != 0 || child_ptr[n_nodes] != n_nodes - 1) return fail(TTB_EINVAL, "ttb_set_tree: child_ptr must cover n_nodes-1 children");
  std::vector<int> height(n_nodes, 0), depth(n_nodes, 0), slot(n_nodes, -1);
  int n_int = 0, n_tips = 0;
  for (int n = 0; n < n_nodes; ++n) {
    const int nc = child_ptr[n + 1] - child_ptr[n];
    if (nc < 0) return fail(TTB_EINVAL, "ttb_set_tree: child_ptr not monotone");
    if (n > 0 && (parent[n] < 0 || parent[n] >= n)) return fail(TTB_EINVAL, "ttb_set_tree: nodes must be in preorder (parent id < child id)");
    if ((nc == 0) != (tip_row[n] >= 0)) return fail(TTB_EINVAL, "ttb_set_tree: tip_row must be >= 0 exactly for nodes without children");
    if (nc == 0) {
      if (tip_row[n] != n_tips) return fail(TTB_EINVAL, "ttb_set_tree: tip rows must be numbered in node order");
      ++n_tips;
    } else {
      slot[n] = n_int++;
    }
    for (int k = child_ptr[n]; k < child_ptr[n + 1]; ++k) {
      const int c = child_idx[k];
      if (c <= n || c >= n_nodes || parent[c] != n) return fail(TTB_EINVAL, "ttb_set_tree: child_idx inconsistent with parent");
    }
    if (n > 0) depth[n] = depth[parent[n]] + 1;
  }
  for (int n = n_nodes - 1; n > 0; --n) height[parent[n]] = std::max(height[parent[n]], height[n] + 1);
  CK(cudaStreamSynchronize(h->stream));
  h->n_nodes = n_nodes;
  h->n_int = n_int;
  h->n_tips = n_tips;
  h->parent.assign(parent, parent + n_nodes);
  h->child_ptr.assign(child_ptr, child_ptr + n_nodes + 1);
  h->child_idx.assign(child_idx, child_idx + n_nodes - 1);
  h->tip_row.assign(tip_row, tip_row + n_nodes);
  h->int_slot = slot;
  std::vector<char> is_int(n_nodes), has_int_child(n_nodes, 0);
  for (int n = 0; n < n_nodes; ++n) is_int[n] = slot[n] >= 0;
  for (int n = 1; n < n_nodes; ++n)
    if (is_int[n]) has_int_child[parent[n]] = 1;
  h->tip_nodes.assign(n_tips, 0);
  for (int n = 0; n < n_nodes; ++n)
    if (tip_row[n] >= 0) h->tip_nodes[tip_row[n]] = n;
  {  // joint back-trace lists: non-root nodes grouped by depth
    int maxd = 0;
    for (int n = 1; n < n_nodes; ++n) maxd = std::max(maxd, depth[n]);
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<int>& nodes = pass ? h->jpre_all_nodes : h->jpre_int_nodes;
      std::vector<TtbLevelLaunch>& levels = pass ? h->jpre_all_levels : h->jpre_int_levels;
      nodes.clear();
      levels.clear();
      std::vector<std::vector<int>> by(maxd + 1);
      for (int n = 1; n < n_nodes; ++n)
        if (pass || slot[n] >= 0) by[depth[n]].push_back(n);
      for (int dd = 1; dd <= maxd; ++dd)
        if (!by[dd].empty()) {
          levels.push_back({(int)nodes.size(), (int)by[dd].size()});
          nodes.insert(nodes.end(), by[dd].begin(), by[dd].end());
        }
    }
  }
  build_sched(h, height, is_int, true, [](int) { return true; }, h->post);
  build_sched(h, depth, has_int_child, false, [&](int c) { return slot[c] >= 0; }, h->pre_int);
  build_sched(h, depth, is_int, false, [](int) { return true; }, h->pre_all);
  h->sched_tiles = -1;
  cudaStream_t s = h->stream;
  int rc;
  if ((rc = upload(h->d_parent, h->parent.data(), h->parent.size(), s))) return rc;
  if ((rc = upload(h->d_child_ptr, h->child_ptr.data(), h->child_ptr.size(), s))) return rc;
  if ((rc = upload(h->d_child_idx, h->child_idx.data(), h->child_idx.size(), s))) return rc;
  if ((rc = upload(h->d_tip_row, h->tip_row.data(), h->tip_row.size(), s))) return rc;
  if ((rc = upload(h->d_int_slot, h->int_slot.data(), h->int_slot.size(), s))) return rc;
  if ((rc = upload(h->d_tip_nodes, h->tip_nodes.data(), h->tip_nodes.size(), s))) return rc;
  if ((rc = upload(h->d_jpre_int, h->jpre_int_nodes.data(), h->jpre_int_nodes.size(), s))) return rc;
  if ((rc = upload(h->d_jpre_all, h->jpre_all_nodes.data(), h->jpre_all_nodes.size(), s))) return rc;
  for (Sched* sc : {&h->post, &h->pre_int, &h->pre_all}) {
    if ((rc = upload(sc->d_chunks, sc->chunks.data(), sc->chunks.size(), s))) return rc;
    if ((rc = upload(sc->d_dep, sc->dep.data(), sc->dep.size(), s))) return rc;
  }
  if ((rc = upload(h->post.d_node_chunk, h->post.node_chunk.data(), h->post.node_chunk.size(), s))) return rc;
  CK(cudaStreamSynchronize(s));
  // every per-node array is invalid now
  h->d_P.release(); h->d_Pf.release(); h->d_TU.release(); h->d_S.release(); h->d_F.release(); h->d_M.release(); h->d_Mtip.release();
  h->d_idx.release(); h->d_idxtip.release();
  h->d_t.release();
  h->have_t = false;
  h->have_pass = h->have_tip_pass = false;
  h->have_joint = h->have_joint_tips = false;
  h->d_LP.release(); h->d_TL.release(); h->d_TC.release(); h->d_Cx.release();
  h->have_masks = false;      // masks are per node: the caller sends them again for the new tree
  h->d_mask_id.release(); h->d_masks.release();
  h->first_full = true;
  h->drop_graphs();
  return 0;
}

}  // extern "C" (templates need C++ linkage)

// Shared by the dense and the sparse setter: sizes, tables, multiplicities, invalidation.
// `fill` uploads the tip codes into h->d_codes (already allocated for the new size).
template <typename Fill>
static int set_patterns_common(ttb_handle h, int64_t n_patterns, int32_t n_codes, const double* code_profiles,
                               const double* multiplicity, Fill fill) {
  if (int rc = use_device(h)) return rc;
  if (!h->n_nodes) return fail(TTB_EINVAL, "ttb_set_patterns: call ttb_set_tree first");
  if (n_patterns <= 0 || n_codes <= 0 || n_codes > 255 || !code_profiles || !multiplicity)
    return fail(TTB_EINVAL, "ttb_set_patterns: bad arguments");
  if ((size_t)n_codes * h->q * 8 > 16 * 1024) return fail(TTB_EUNSUPPORTED, "ttb_set_patterns: too many distinct characters for the tip tables");
  const long long Lp = n_patterns, ld = (Lp + 31) / 32 * 32;
  // same sizes as before: nothing is reallocated and the call is stream-ordered without host syncs
  if (ld != h->ld || Lp != h->Lp || n_codes != h->n_codes) CK(cudaStreamSynchronize(h->stream));
  cudaStream_t s = h->stream;
  int rc;
  if ((rc = h->d_codes.alloc((size_t)h->n_tips * ld))) return rc;
  if ((rc = fill(Lp, ld, s))) return rc;
  if ((rc = h->d_code_prof.alloc((size_t)n_codes * h->q))) return rc;
  if ((rc = upload_small(h, h->d_code_prof.p, code_profiles, (size_t)n_codes * h->q * sizeof(double)))) return rc;
  if ((rc = h->d_mult.alloc(ld))) return rc;
  CK(cudaMemsetAsync(h->d_mult.p, 0, h->d_mult.bytes(), s));
  if ((rc = upload_small(h, h->d_mult.p, multiplicity, Lp * sizeof(double)))) return rc;
  h->h_mult.assign(multiplicity, multiplicity + Lp);
  if (ld != h->ld || Lp != h->Lp) {
    if (h->site_specific) { h->site_specific = false; h->have_gtr = false; }   // per-pattern model no longer matches
    h->d_S.release(); h->d_F.release(); h->d_M.release(); h->d_Mtip.release();
    h->d_idx.release(); h->d_idxtip.release(); h->d_LH.release(); h->d_lh_partial.release();
    h->drop_graphs();
  } else if (n_codes != h->n_codes) {
    h->d_TU.release();
    h->drop_graphs();  // n_codes is a kernel parameter
  }
  if (n_codes != h->n_codes) h->prepared = false;   // the stages' shared-memory size depends on the code table
  if (h->d_idx.p) CK(cudaMemsetAsync(h->d_idx.p, 0xff, h->d_idx.bytes(), s));  // new data: no previous states
  if (h->d_idxtip.p) CK(cudaMemsetAsync(h->d_idxtip.p, 0xff, h->d_idxtip.bytes(), s));
  h->Lp = Lp;
  h->ld = ld;
  h->n_codes = n_codes;
  h->have_pass = h->have_tip_pass = false;
  if (h->have_masks) {        // masks are per pattern: the caller sends them again for the new alignment
    h->have_masks = false;
    h->d_mask_id.release(); h->d_masks.release();
    h->drop_graphs();
  }
  h->first_full = true;
  return 0;
}

extern "C" {

int ttb_set_patterns(ttb_handle h, int64_t n_patterns, const uint8_t* tip_codes, int32_t n_codes,
                     const double* code_profiles, const double* multiplicity) {
  if (!h) return fail(TTB_EINVAL, "null handle");
  if (!tip_codes) return fail(TTB_EINVAL, "ttb_set_patterns: bad arguments");
  return set_patterns_common(h, n_patterns, n_codes, code_profiles, multiplicity, [&](long long Lp, long long ld, cudaStream_t s) -> int {
    // one contiguous H2D copy into a packed staging buffer, re-pitched to the padded layout on the device
    if (int rc = h->d_bstage.alloc(std::max((size_t)h->n_tips, (size_t)h->n_int) * (size_t)Lp)) return rc;
    CK(cudaMemcpyAsync(h->d_bstage.p, tip_codes, (size_t)h->n_tips * Lp, cudaMemcpyHostToDevice, s));
    pitch_bytes_kernel<<<h->n_sm * 8, 256, 0, s>>>(h->d_bstage.p, Lp, h->d_codes.p, ld, Lp, h->n_tips, 0);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
  });
}

int ttb_alignment_stats(ttb_handle h, int64_t n_seq, int64_t L, const uint8_t* aln, int32_t fill_overhangs, int32_t gap,
                        int32_t fill, int32_t ambiguous, uint8_t* lo, uint8_t* hi, uint8_t* all_amb) {
  if (int rc = use_device(h)) return rc;
  if (n_seq <= 0 || L <= 0 || !aln || !lo || !hi || !all_amb) return fail(TTB_EINVAL, "ttb_alignment_stats: bad arguments");
  cudaStream_t s = h->stream;
  int rc;
  if ((rc = h->d_aln.alloc((size_t)n_seq * L))) return rc;
  if ((rc = h->d_colstat.alloc((size_t)3 * L))) return rc;
  CK(cudaMemcpyAsync(h->d_aln.p, aln, (size_t)n_seq * L, cudaMemcpyHostToDevice, s));
  if (fill_overhangs) {
    fill_overhangs_kernel<<<(unsigned)((n_seq + 3) / 4), 128, 0, s>>>(h->d_aln.p, L, n_seq, (uint8_t)gap, (uint8_t)fill);
    h->launches += 1;
  }
  column_stats_kernel<<<(unsigned)((L + 127) / 128), 128, 0, s>>>(h->d_aln.p, L, n_seq, ambiguous, h->d_colstat.p, h->d_colstat.p + L,
                                                                h->d_colstat.p + 2 * L);
  h->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(lo, h->d_colstat.p, (size_t)L, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hi, h->d_colstat.p + L, (size_t)L, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(all_amb, h->d_colstat.p + 2 * L, (size_t)L, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  h->aln_rows = n_seq;
  h->aln_L = L;
  return 0;
}

int ttb_set_patterns_from_alignment(ttb_handle h, int64_t n_patterns, const int64_t* first_pos, const uint8_t* const_letter,
                                    const int32_t* tip_seq_row, const uint8_t* lut, int32_t missing_code, int32_t n_codes,
                                    const double* code_profiles, const double* multiplicity) {
  if (!h) return fail(TTB_EINVAL, "null handle");
  if (!h->d_aln.p) return fail(TTB_EINVAL, "ttb_set_patterns_from_alignment: call ttb_alignment_stats first");
  if (!first_pos || !const_letter || !tip_seq_row || !lut || missing_code < 0 || missing_code >= n_codes)
    return fail(TTB_EINVAL, "ttb_set_patterns_from_alignment: bad arguments");
  for (int64_t i = 0; i < n_patterns; ++i)
    if (!const_letter[i] && (first_pos[i] < 0 || first_pos[i] >= h->aln_L)) return fail(TTB_EINVAL, "ttb_set_patterns_from_alignment: column out of range");
  for (int t = 0; t < h->n_tips; ++t)
    if (tip_seq_row[t] >= h->aln_rows) return fail(TTB_EINVAL, "ttb_set_patterns_from_alignment: alignment row out of range");
  int bad = 0;
  int rc0 = set_patterns_common(h, n_patterns, n_codes, code_profiles, multiplicity, [&](long long Lp, long long ld, cudaStream_t s) -> int {
    int rc;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    if ((rc = upload(h->d_firstpos, reinterpret_cast<const long long*>(first_pos), (size_t)Lp, s))) return rc;
    if ((rc = upload(h->d_constl, const_letter, (size_t)Lp, s))) return rc;
    if ((rc = upload(h->d_seqrow, tip_seq_row, (size_t)h->n_tips, s))) return rc;
    if ((rc = upload(h->d_lut, lut, (size_t)256, s))) return rc;
    if ((rc = h->d_flag.alloc(1))) return rc;
    CK(cudaMemsetAsync(h->d_flag.p, 0, sizeof(int), s));
    gather_patterns_kernel<<<h->n_sm * 8, 256, 0, s>>>(h->d_aln.p, h->aln_L, h->d_firstpos.p, h->d_constl.p, h->d_seqrow.p, h->d_lut.p,
                                                    missing_code, Lp, ld, h->n_tips, h->d_codes.p, h->d_flag.p);
    h->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&bad, h->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
  });
  if (rc0) return rc0;
  if (bad) {
    char buf[128];
    snprintf(buf, sizeof buf, "ttb_set_patterns_from_alignment: character '%c' (%d) is not in the profile map", (char)(bad - 1), bad - 1);
    return fail(TTB_EINVAL, buf);
  }
  return 0;
}

int ttb_set_patterns_sparse(ttb_handle h, int64_t n_patterns, const uint8_t* ref_codes, int64_t n_entries,
                            const int32_t* entry_row, const int32_t* entry_pos, const uint8_t* entry_code,
                            int32_t n_codes, const double* code_profiles, const double* multiplicity) {
  if (!h) return fail(TTB_EINVAL, "null handle");
  if (!ref_codes || n_entries < 0 || (n_entries && (!entry_row || !entry_pos || !entry_code)))
    return fail(TTB_EINVAL, "ttb_set_patterns_sparse: bad arguments");
  for (int64_t i = 0; i < n_entries; ++i)
    if (entry_row[i] < 0 || entry_row[i] >= h->n_tips || entry_pos[i] < 0 || entry_pos[i] >= n_patterns || entry_code[i] >= n_codes)
      return fail(TTB_EINVAL, "ttb_set_patterns_sparse: entry out of range");
  return set_patterns_common(h, n_patterns, n_codes, code_profiles, multiplicity, [&](long long Lp, long long ld, cudaStream_t s) -> int {
    int rc;
    if ((rc = h->d_bstage.alloc(std::max((size_t)h->n_tips, (size_t)h->n_int) * (size_t)Lp))) return rc;
    CK(cudaMemcpyAsync(h->d_bstage.p, ref_codes, (size_t)Lp, cudaMemcpyHostToDevice, s));
    fill_ref_codes_kernel<<<h->n_sm * 8, 256, 0, s>>>(h->d_bstage.p, h->d_codes.p, ld, Lp, h->n_tips);
    h->launches += 1;
    if (n_entries) {
      if ((rc = upload(h->d_ent_row, entry_row, (size_t)n_entries, s))) return rc;
      if ((rc = upload(h->d_ent_pos, entry_pos, (size_t)n_entries, s))) return rc;
      if ((rc = h->d_mut_state.alloc(std::max(h->d_mut_state.n, (size_t)n_entries)))) return rc;
      CK(cudaMemcpyAsync(h->d_mut_state.p, entry_code, (size_t)n_entries, cudaMemcpyHostToDevice, s));
      scatter_codes_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, s>>>(h->d_ent_row.p, h->d_ent_pos.p, h->d_mut_state.p,
                                                                               n_entries, h->d_codes.p, ld);
      h->launches += 1;
    }
    CK(cudaGetLastError());
    return 0;
  });
}

int ttb_set_gtr(ttb_handle h, const double* eigvals, const double* v, const double* v_inv, const double* Pi,
                double mu, int32_t gap_index) {
  if (int rc = use_device(h)) return rc;
  if (!eigvals || !v || !v_inv || !Pi) return fail(TTB_EINVAL, "ttb_set_gtr: null argument");
  const size_t q = h->q;
  cudaStream_t s = h->stream;
  int rc;
  const bool realloc = !h->d_eig.p;
  (void)s;
  if ((rc = h->d_eig.alloc(q)) || (rc = h->d_v.alloc(q * q)) || (rc = h->d_vinv.alloc(q * q)) || (rc = h->d_Pi.alloc(q)) ||
      (rc = h->d_mu.alloc(1)))
    return rc;
  if ((rc = upload_small(h, h->d_eig.p, eigvals, q * sizeof(double)))) return rc;
  if ((rc = upload_small(h, h->d_v.p, v, q * q * sizeof(double)))) return rc;
  if ((rc = upload_small(h, h->d_vinv.p, v_inv, q * q * sizeof(double)))) return rc;
  if ((rc = upload_small(h, h->d_Pi.p, Pi, q * sizeof(double)))) return rc;
  if ((rc = upload_small(h, h->d_mu.p, &mu, sizeof(double)))) return rc;
  if (realloc || h->site_specific) h->drop_graphs();
  h->site_specific = false;
  h->mu = mu;
  h->gap_index = gap_index;
  h->have_gtr = true;
  return 0;
}

// Upload a [rows][Lp] host plane set into a [rows][ld] device buffer (zero padded).
static int upload_planes(ttb_handle h, DBuf<double>& b, const double* src, size_t rows) {
  if (int rc = b.alloc(rows * (size_t)h->ld)) return rc;
  CK(cudaMemsetAsync(b.p, 0, b.bytes(), h->stream));
  CK(cudaMemcpy2DAsync(b.p, h->ld * sizeof(double), src, h->Lp * sizeof(double), h->Lp * sizeof(double), rows,
                       cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int ttb_set_gtr_site_specific(ttb_handle h, const double* eigvals, const double* v, const double* v_inv, const double* Pi,
                              const double* mu, const double* t_grid, int32_t n_grid, double rate_scale,
                              int32_t approximate, int32_t gap_index) {
  if (int rc = use_device(h)) return rc;
  if (!eigvals || !v || !v_inv || !Pi || !mu || !t_grid || n_grid < 2) return fail(TTB_EINVAL, "ttb_set_gtr_site_specific: bad arguments");
  if (h->q > 8) return fail(TTB_EUNSUPPORTED, "site-specific models are compiled for alphabets of up to 8 states");
  if (!h->Lp) return fail(TTB_EINVAL, "ttb_set_gtr_site_specific: call ttb_set_patterns first (the model is per pattern)");
  const size_t q = h->q, L = (size_t)h->Lp;
  // reference layout (gtr_site_specific.py:327-329): v[k][i][a] = V_a[i][k], v_inv[j][k][a] = Vinv_a[k][j];
  // device planes: V at (i*q+k), Vinv at (k*q+j)  => transpose the two leading axes on the host
  std::vector<double> V(q * q * L), Vi(q * q * L);
  for (size_t i = 0; i < q; ++i)
    for (size_t k = 0; k < q; ++k) {
      memcpy(&V[(i * q + k) * L], &v[(k * q + i) * L], L * sizeof(double));
      memcpy(&Vi[(k * q + i) * L], &v_inv[(i * q + k) * L], L * sizeof(double));
    }
  int rc;
  if ((rc = upload_planes(h, h->d_ss_eig, eigvals, q))) return rc;
  if ((rc = upload_planes(h, h->d_ss_mu, mu, 1))) return rc;
  if ((rc = upload_planes(h, h->d_ss_V, V.data(), q * q))) return rc;
  if ((rc = upload_planes(h, h->d_ss_Vinv, Vi.data(), q * q))) return rc;
  if ((rc = upload_planes(h, h->d_ss_Pi, Pi, q))) return rc;
  if ((rc = upload(h->d_ss_grid, t_grid, (size_t)n_grid, h->stream))) return rc;
  h->ss_grid.assign(t_grid, t_grid + n_grid);
  if ((rc = h->d_ss_E.alloc((size_t)h->tiles() * n_grid * q * TTB_TILE))) return rc;   // tile-blocked, see TtbDev::ss_E
  {
    const bool was = h->site_specific;
    h->site_specific = true;
    ss_grid_table_kernel<<<h->n_sm * 8, 256, 0, h->stream>>>(h->dev(), h->d_ss_E.p);
    h->site_specific = was;
    h->launches += 1;
    CK(cudaGetLastError());
  }
  // Symmetric structure of a reversible model's eigen-system (GTR._eig_single_site, gtr.py:612-629):
  // Vinv[k][j] = V[j][k] * c_k / Pi[j].  If every pattern has it (to 1e-9 of the largest entry) the level kernels use
  // the symmetric form; anything else keeps the general kernels.  TTB_SS_SYM=0 in the environment forces the latter.
  bool sym = q <= TTB_SS_REG_MAXQ;
  if (const char* e = getenv("TTB_SS_SYM")) sym = sym && atoi(e) != 0;
  std::vector<double> C;
  if (sym) {
    C.assign(q * L, 0.0);
    for (size_t a = 0; a < L && sym; ++a) {
      double vmax = 0.0;
      for (size_t r = 0; r < q * q; ++r) vmax = std::max(vmax, std::fabs(Vi[r * L + a]));
      for (size_t k = 0; k < q && sym; ++k) {
        size_t js = 0;
        for (size_t j = 1; j < q; ++j)
          if (std::fabs(V[(j * q + k) * L + a]) > std::fabs(V[(js * q + k) * L + a])) js = j;
        const double vjk = V[(js * q + k) * L + a];
        if (vjk == 0.0) { sym = false; break; }
        const double c = Vi[(k * q + js) * L + a] * Pi[js * L + a] / vjk;
        C[k * L + a] = c;
        for (size_t j = 0; j < q; ++j) {
          const double pj = Pi[j * L + a];
          if (!(pj > 0.0) || !(std::fabs(Vi[(k * q + j) * L + a] - V[(j * q + k) * L + a] * c / pj) <= 1e-9 * vmax)) { sym = false; break; }
        }
      }
    }
  }
  h->ss_sym = sym;
  if (sym) {
    if ((rc = upload_planes(h, h->d_ss_c, C.data(), q))) return rc;
    if ((rc = h->d_ss_Ec.alloc((size_t)h->tiles() * n_grid * q * TTB_TILE))) return rc;
    const bool was = h->site_specific;
    h->site_specific = true;
    ss_grid_table_kernel<<<h->n_sm * 8, 256, 0, h->stream>>>(h->dev(), h->d_ss_Ec.p, h->d_ss_c.p);
    h->site_specific = was;
    h->launches += 1;
    CK(cudaGetLastError());
  } else {
    h->d_ss_c.release();
    h->d_ss_Ec.release();
  }
  CK(cudaStreamSynchronize(h->stream));   // V / Vi / C are locals
  h->ss_tmax = approximate ? 10.0 / rate_scale : 0.0;
  h->ss_interp_dirty = true;
  h->gap_index = gap_index;
  h->site_specific = true;
  h->have_gtr = true;
  h->drop_graphs();
  return 0;
}

// Bracket every branch length on the interpolation grid exactly like scipy's interp1d(kind='linear',
// assume_sorted=True): idx = searchsorted(grid, t) clipped to [1, n-1]  (gtr_site_specific.py:345-348,367-371).
static int update_ss_interp(ttb_handle h) {
  if (!h->site_specific || !h->ss_interp_dirty) return 0;
  const int n = h->n_nodes, ng = (int)h->ss_grid.size();
  std::vector<double> w(n, -1.0);
  std::vector<int> glo(n, 0);
  for (int i = 1; i < n; ++i) {
    const double t = h->h_t[i];
    if (h->ss_tmax > 0.0 && t < h->ss_tmax) {
      int lo = (int)(std::lower_bound(h->ss_grid.begin(), h->ss_grid.end(), t) - h->ss_grid.begin());
      lo = std::max(1, std::min(ng - 1, lo));
      glo[i] = lo - 1;
      w[i] = (t - h->ss_grid[lo - 1]) / (h->ss_grid[lo] - h->ss_grid[lo - 1]);
    }
  }
  std::vector<double2> rec(n);
  for (int i = 0; i < n; ++i) rec[i] = make_double2(w[i], i ? h->h_t[i] : 0.0);
  const bool realloc = !h->d_ss_w.p;
  int rc;
  if ((rc = upload(h->d_ss_lo, glo.data(), (size_t)n, h->stream))) return rc;
  if ((rc = upload(h->d_ss_w, w.data(), (size_t)n, h->stream))) return rc;
  if ((rc = upload(h->d_ss_rec, rec.data(), (size_t)n, h->stream))) return rc;
  // the producer warps read each child's bracket from its chunk descriptor
  for (Sched* sc : {&h->post, &h->pre_int, &h->pre_all}) {
    const int nc = (int)sc->chunks.size();
    if (nc && sc->d_chunks.p) {
      ss_patch_chunks_kernel<<<(nc + 255) / 256, 256, 0, h->stream>>>(sc->d_chunks.p, nc, h->d_ss_lo.p);
      h->launches += 1;
    }
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (realloc) h->drop_graphs();
  h->ss_interp_dirty = false;
  return 0;
}

int ttb_set_branch_masks(ttb_handle h, int32_t n_masks, const uint8_t* masks, const int32_t* node_mask) {
  if (int rc = use_device(h)) return rc;
  if (!h->n_nodes || !h->Lp) return fail(TTB_EINVAL, "ttb_set_branch_masks: call ttb_set_tree and ttb_set_patterns first");
  const bool had = h->have_masks;
  if (n_masks <= 0 || !masks || !node_mask) {
    h->have_masks = false;
    h->d_mask_id.release();
    h->d_masks.release();
    if (had) h->drop_graphs();
    return 0;
  }
  for (int n = 0; n < h->n_nodes; ++n)
    if (node_mask[n] < -1 || node_mask[n] >= n_masks) return fail(TTB_EINVAL, "ttb_set_branch_masks: mask index out of range");
  const size_t Lp = (size_t)h->Lp, ld = (size_t)h->ld;
  for (size_t i = 0; i < (size_t)n_masks * Lp; ++i)
    if (masks[i] > 1) return fail(TTB_EINVAL, "ttb_set_branch_masks: masks must be 0 or 1");
  int rc;
  if ((rc = h->d_masks.alloc((size_t)n_masks * ld))) return rc;
  CK(cudaMemsetAsync(h->d_masks.p, 0, h->d_masks.bytes(), h->stream));
  CK(cudaMemcpy2DAsync(h->d_masks.p, ld, masks, Lp, Lp, (size_t)n_masks, cudaMemcpyHostToDevice, h->stream));
  if ((rc = upload(h->d_mask_id, node_mask, (size_t)h->n_nodes, h->stream))) return rc;
  CK(cudaStreamSynchronize(h->stream));   // pageable sources
  h->have_masks = true;
  h->drop_graphs();     // other kernels, other arguments
  return 0;
}

int ttb_set_branch_lengths(ttb_handle h, const double* t) {
  if (int rc = use_device(h)) return rc;
  if (!h->n_nodes || !t) return fail(TTB_EINVAL, "ttb_set_branch_lengths: call ttb_set_tree first");
  const bool realloc = !h->d_t.p;
  if (int rc = h->d_t.alloc(h->n_nodes)) return rc;
  // pageable source: the copy is staged before the call returns, so the caller may reuse `t`
  if (int rc = upload_small(h, h->d_t.p, t, h->n_nodes * sizeof(double))) return rc;
  h->h_t.assign(t, t + h->n_nodes);
  h->ss_interp_dirty = true;
  if (realloc) h->drop_graphs();
  h->have_t = true;
  return 0;
}

int ttb_marginal(ttb_handle h, int32_t flags) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, false)) return rc;
  if (h->f32 && h->have_masks) return fail(TTB_EUNSUPPORTED, "float message storage is not available together with per-branch masks");
  const bool lh_only = flags & TTB_LH_ONLY;
  const bool tips = (flags & TTB_RECONSTRUCT_TIPS) && !lh_only;
  const bool keep_prev = (flags & TTB_KEEP_PREV_STATES) && !lh_only;
  flags = (lh_only ? TTB_LH_ONLY : 0) | (tips ? TTB_RECONSTRUCT_TIPS : 0);
  const bool had_P = h->d_P.p != nullptr;
  if (int rc = ensure_state(h, tips)) return rc;
  if (!had_P) h->drop_graphs();
  if (!lh_only)
    if (int rc = ensure_preorder_state(h, tips)) return rc;
  if (int rc = update_ss_interp(h)) return rc;
  if (keep_prev) {   // the states this pass is about to overwrite: what ttb_sample_states counts changes against
    if (int rc = h->d_idx_prev.alloc(h->d_idx.n)) return rc;
    CK(cudaMemcpyAsync(h->d_idx_prev.p, h->d_idx.p, h->d_idx.bytes(), cudaMemcpyDeviceToDevice, h->stream));
    if (tips) {
      if (int rc = h->d_idxtip_prev.alloc(h->d_idxtip.n)) return rc;
      CK(cudaMemcpyAsync(h->d_idxtip_prev.p, h->d_idxtip.p, h->d_idxtip.bytes(), cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  if (!lh_only) {
    h->have_prev = keep_prev;
    h->have_prev_tips = keep_prev && tips;
  }
  const int count_diff = lh_only ? 0 : 1;
  const int key = flags;
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    // capture on the engine's own stream, replay on whatever stream is current
    cudaGraph_t graph;
    int nk = 0;
    CK(cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_pass(h, flags, count_diff, h->own_stream, &nk);
    cudaError_t ce = cudaStreamEndCapture(h->own_stream, &graph);
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(TTB_ECUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    cudaGraphExec_t exec;
    CK(cudaGraphInstantiate(&exec, graph, 0));
    CK(cudaGraphDestroy(graph));
    h->graphs[key] = exec;
    h->graph_kernels[key] = nk;
    it = h->graphs.find(key);
  }
  CK(cudaGraphLaunch(it->second, h->stream));
  CK(cudaMemcpyAsync(h->h_results, h->d_results.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->launches += h->graph_kernels[key];
  if (!lh_only) {
    h->have_pass = true;
    h->have_tip_pass = tips;
  }
  h->have_joint = h->have_joint_tips = false;
  return 0;
}

int ttb_joint(ttb_handle h, int32_t flags) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, false)) return rc;
  if (h->site_specific) return fail(TTB_EUNSUPPORTED, "ttb_joint: joint reconstruction is not implemented for site-specific models");
  if (h->f32) return fail(TTB_EUNSUPPORTED, "ttb_joint: the joint pass keeps log-space sums in double; call ttb_set_message_storage(TTB_STORAGE_F64) first");
  if (h->have_masks) return fail(TTB_EUNSUPPORTED, "ttb_joint: joint reconstruction is not implemented with per-branch masks");
  const bool tips = flags & TTB_RECONSTRUCT_TIPS;
  const bool trace = !(flags & TTB_JOINT_NO_TRACE);
  const bool had_P = h->d_P.p != nullptr;
  if (int rc = ensure_state(h, tips)) return rc;
  if (!had_P) h->drop_graphs();
  const size_t q = h->q;
  const size_t pq = (q * q + 1) / 2 * 2, tus = ((size_t)h->n_codes * q + 1) / 2 * 2;
  int rc;
  const bool fresh = !h->d_Cx.p || !h->d_idx.p || (tips && !h->d_idxtip.p);
  if ((rc = h->d_LP.alloc((size_t)h->n_nodes * pq))) return rc;
  if ((rc = h->d_TL.alloc((size_t)h->n_tips * tus))) return rc;
  if ((rc = h->d_TC.alloc((size_t)h->n_tips * tus))) return rc;
  if ((rc = h->d_Cx.alloc((size_t)h->n_int * h->tiles() * q * TTB_TILE))) return rc;
  if (!h->d_idx.p) {
    if ((rc = h->d_idx.alloc((size_t)h->n_int * h->ld))) return rc;
    CK(cudaMemsetAsync(h->d_idx.p, 0xff, h->d_idx.bytes(), h->stream));
  }
  if (tips && !h->d_idxtip.p) {
    if ((rc = h->d_idxtip.alloc((size_t)h->n_tips * h->ld))) return rc;
    CK(cudaMemsetAsync(h->d_idxtip.p, 0xff, h->d_idxtip.bytes(), h->stream));
  }
  if (fresh) h->drop_graphs();
  const int key = 8 | (tips ? TTB_RECONSTRUCT_TIPS : 0) | (trace ? 0 : 16);
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    TtbPassPlan pl;
    fill_plan(h, pl, tips, 1);
    cudaGraph_t graph;
    CK(cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeThreadLocal));
    const int nk = ttb_qops(h->q)->enqueue_joint(pl, h->own_stream, trace ? 1 : 0);
    cudaError_t ce = cudaStreamEndCapture(h->own_stream, &graph);
    if (ce != cudaSuccess) return fail(TTB_ECUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    if (nk <= 0) return fail(TTB_EUNSUPPORTED, "ttb_joint: not available for this model");
    cudaGraphExec_t exec;
    CK(cudaGraphInstantiate(&exec, graph, 0));
    CK(cudaGraphDestroy(graph));
    h->graphs[key] = exec;
    h->graph_kernels[key] = nk;
    it = h->graphs.find(key);
  }
  CK(cudaGraphLaunch(it->second, h->stream));
  CK(cudaMemcpyAsync(h->h_results, h->d_results.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->launches += h->graph_kernels[key];
  h->have_pass = h->have_tip_pass = false;   // the marginal messages were overwritten
  h->have_joint = true;
  h->have_joint_tips = tips && trace;
  return 0;
}

int ttb_joint_retrace(ttb_handle h, const uint8_t* root_idx, int32_t flags) {
  if (int rc = use_device(h)) return rc;
  if (!h->have_joint) return fail(TTB_EINVAL, "ttb_joint_retrace: call ttb_joint first");
  if (!root_idx) return fail(TTB_EINVAL, "ttb_joint_retrace: null root_idx");
  const bool tips = flags & TTB_RECONSTRUCT_TIPS;
  if (tips && !h->d_idxtip.p) return fail(TTB_EINVAL, "ttb_joint_retrace: ttb_joint was run without TTB_RECONSTRUCT_TIPS");
  for (long long a = 0; a < h->Lp; ++a)
    if (root_idx[a] >= h->q) return fail(TTB_EINVAL, "ttb_joint_retrace: root state out of range");
  int rc;
  if ((rc = h->d_bstage.alloc(std::max((size_t)h->n_tips, (size_t)h->n_int) * (size_t)h->Lp))) return rc;
  CK(cudaMemcpyAsync(h->d_bstage.p, root_idx, (size_t)h->Lp, cudaMemcpyHostToDevice, h->stream));
  TtbPassPlan pl;
  fill_plan(h, pl, tips, 1);
  const int nk = ttb_qops(h->q)->enqueue_joint_retrace(pl, h->d_bstage.p, h->stream);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_results, h->d_results.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->launches += nk;
  h->have_joint_tips = tips;
  return 0;
}

int ttb_sample_states(ttb_handle h, int32_t n, const int32_t* nodes, const double* uniforms, int64_t* n_diff,
                      int64_t* n_diff_tips) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, true)) return rc;
  if (n < 0 || (n && (!nodes || !uniforms))) return fail(TTB_EINVAL, "ttb_sample_states: bad arguments");
  if (!h->have_prev) return fail(TTB_EINVAL, "ttb_sample_states: the last ttb_marginal was not run with TTB_KEEP_PREV_STATES");
  for (int k = 0; k < n; ++k) {
    const int node = nodes[k];
    if (node < 0 || node >= h->n_nodes) return fail(TTB_EINVAL, "ttb_sample_states: bad node id");
    if (h->tip_row[node] >= 0 && !(h->have_tip_pass && h->have_prev_tips))
      return fail(TTB_EINVAL, "ttb_sample_states: tip profiles exist only after TTB_RECONSTRUCT_TIPS");
  }
  cudaStream_t s = h->stream;
  int rc;
  if ((rc = h->d_scount.alloc(2))) return rc;
  CK(cudaMemsetAsync(h->d_scount.p, 0, 2 * sizeof(unsigned long long), s));
  const size_t Lp = (size_t)h->Lp;
  // the uniforms travel in blocks of at most 256 MB (TTB_SAMPLE_BLOCK_DOUBLES overrides the block size: tests)
  size_t blk_doubles = (size_t)1 << 25;
  if (const char* e = getenv("TTB_SAMPLE_BLOCK_DOUBLES")) blk_doubles = (size_t)std::max(1LL, atoll(e));
  const int blk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(n, 1), blk_doubles / Lp, (size_t)65535}));   // grid.y <= 65535
  if (n) {
    if ((rc = h->d_sg_uniforms.alloc((size_t)blk * Lp))) return rc;
    if ((rc = h->d_enodes.alloc((size_t)n))) return rc;
    CK(cudaMemcpyAsync(h->d_enodes.p, nodes, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
  }
  for (int k0 = 0; k0 < n; k0 += blk) {
    const int m = std::min(blk, n - k0);
    CK(cudaMemcpyAsync(h->d_sg_uniforms.p, uniforms + (size_t)k0 * Lp, (size_t)m * Lp * sizeof(double), cudaMemcpyHostToDevice, s));
    ttb_qops(h->q)->sample_states(h->dev(), h->tiles(), m, h->d_enodes.p + k0, h->d_sg_uniforms.p, h->d_idx_prev.p,
                                  h->d_idxtip_prev.p, h->d_scount.p, s);
    CK(cudaGetLastError());
    h->launches += 1;
  }
  unsigned long long c[2] = {0, 0};
  CK(cudaMemcpyAsync(c, h->d_scount.p, sizeof c, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  h->d_sg_uniforms.release();
  if (n_diff) *n_diff = (int64_t)c[0];
  if (n_diff_tips) *n_diff_tips = (int64_t)c[1];
  return 0;
}

int ttb_results(ttb_handle h, double* total_lh, int64_t* n_diff) {
  if (int rc = use_device(h)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  if (total_lh) *total_lh = h->h_results[0];
  if (n_diff) *n_diff = (int64_t)h->h_results[1];
  return 0;
}

int ttb_results_tips(ttb_handle h, int64_t* n_diff_tips) {
  if (int rc = use_device(h)) return rc;
  if (!n_diff_tips) return fail(TTB_EINVAL, "null argument");
  CK(cudaStreamSynchronize(h->stream));
  *n_diff_tips = (int64_t)h->h_results[2];
  return 0;
}

int ttb_results_device_ptr(ttb_handle h, void** dptr) {
  if (int rc = use_device(h)) return rc;
  if (!dptr) return fail(TTB_EINVAL, "dptr is null");
  if (int rc = h->d_results.alloc(4)) return rc;
  *dptr = h->d_results.p;
  return 0;
}

int ttb_sync(ttb_handle h) {
  if (int rc = use_device(h)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int ttb_fetch_site_lh(ttb_handle h, double* out) {
  if (int rc = use_device(h)) return rc;
  if (!h->d_LH.p || !out) return fail(TTB_EINVAL, "ttb_fetch_site_lh: nothing computed yet");
  CK(cudaMemcpyAsync(out, h->d_LH.p, h->Lp * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int ttb_fetch_node(ttb_handle h, int32_t node, int32_t which, double* out) {
  if (int rc = use_device(h)) return rc;
  if (node < 0 || node >= h->n_nodes || !out || which < 0 || which > 3) return fail(TTB_EINVAL, "ttb_fetch_node: bad arguments");
  const bool tip = h->tip_row[node] >= 0;
  if (which == TTB_JOINT_ROOT_LX) {
    if (node != 0 || !h->have_joint) return fail(TTB_EINVAL, "ttb_fetch_node: TTB_JOINT_ROOT_LX needs node 0 after ttb_joint");
  } else if (h->have_joint && !h->have_pass) {
    return fail(TTB_EINVAL, "ttb_fetch_node: the marginal messages were overwritten by ttb_joint; run ttb_marginal again");
  } else if (which == TTB_SUBTREE) {
    if (!tip && !h->d_S.p) return fail(TTB_EINVAL, "ttb_fetch_node: run ttb_marginal first");
  } else {
    if (int rc = check_ready(h, true)) return rc;
    if (which == TTB_PROFILE && tip && !h->have_tip_pass)
      return fail(TTB_EINVAL, "ttb_fetch_node: tip profiles exist only after TTB_RECONSTRUCT_TIPS");
  }
  if (int rc = h->d_stage.alloc((size_t)h->Lp * h->q)) return rc;
  ttb_qops(h->q)->fetch_node(h->dev(), h->tiles(), node, which, h->d_stage.p, h->stream);
  h->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->d_stage.p, (size_t)h->Lp * h->q * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int ttb_fetch_seq_idx(ttb_handle h, int32_t n, const int32_t* nodes, uint8_t* out) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_states(h)) return rc;
  if (n < 0 || (n && (!nodes || !out))) return fail(TTB_EINVAL, "ttb_fetch_seq_idx: bad arguments");
  for (int k = 0; k < n; ++k) {
    const int node = nodes[k];
    if (node < 0 || node >= h->n_nodes) return fail(TTB_EINVAL, "ttb_fetch_seq_idx: bad node id");
    const uint8_t* src;
    if (h->tip_row[node] >= 0) {
      if (!h->have_tip_pass && !h->have_joint_tips) return fail(TTB_EINVAL, "ttb_fetch_seq_idx: tip states exist only after TTB_RECONSTRUCT_TIPS");
      src = h->d_idxtip.p + (size_t)h->tip_row[node] * h->ld;
    } else {
      src = h->d_idx.p + (size_t)h->int_slot[node] * h->ld;
    }
    CK(cudaMemcpyAsync(out + (size_t)k * h->Lp, src, h->Lp, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int ttb_fetch_mutations(ttb_handle h, uint8_t* root_idx, int32_t max_n, int32_t* node, int32_t* pos, uint8_t* state,
                        int64_t* n) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_states(h)) return rc;
  if (!root_idx || !n || max_n < 0 || (max_n && (!node || !pos || !state))) return fail(TTB_EINVAL, "ttb_fetch_mutations: bad arguments");
  int rc;
  if ((rc = h->d_mut_node.alloc(std::max(h->d_mut_node.n, (size_t)std::max(max_n, 1))))) return rc;
  if ((rc = h->d_mut_pos.alloc(std::max(h->d_mut_pos.n, (size_t)std::max(max_n, 1))))) return rc;
  if ((rc = h->d_mut_state.alloc(std::max(h->d_mut_state.n, (size_t)std::max(max_n, 1))))) return rc;
  if ((rc = h->d_mut_count.alloc(1))) return rc;
  if ((rc = h->d_mut_offsets.alloc((size_t)h->n_nodes))) return rc;
  cudaStream_t s = h->stream;
  if (h->n_nodes > 1) {
    mut_count_kernel<<<h->n_nodes - 1, 256, 0, s>>>(h->dev(), h->d_mut_offsets.p);
    mut_scan_kernel<<<1, 1024, 0, s>>>(h->d_mut_offsets.p, h->n_nodes, h->d_mut_count.p);
    mut_write_kernel<<<h->n_nodes - 1, 256, 0, s>>>(h->dev(), h->d_mut_offsets.p, (long long)max_n, h->d_mut_node.p, h->d_mut_pos.p, h->d_mut_state.p);
    h->launches += 3;
  } else
    CK(cudaMemsetAsync(h->d_mut_count.p, 0, sizeof(unsigned long long), s));
  CK(cudaGetLastError());
  unsigned long long cnt = 0;
  CK(cudaMemcpyAsync(&cnt, h->d_mut_count.p, sizeof cnt, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(root_idx, h->d_idx.p + (size_t)h->int_slot[0] * h->ld, (size_t)h->Lp, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const size_t m = (size_t)std::min<unsigned long long>(cnt, (unsigned long long)max_n);
  if (m) {
    CK(cudaMemcpyAsync(node, h->d_mut_node.p, m * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(pos, h->d_mut_pos.p, m * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(state, h->d_mut_state.p, m, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  *n = (int64_t)cnt;
  return 0;
}

int ttb_enqueue_fetch_site_lh(ttb_handle h, double* out) {
  if (int rc = use_device(h)) return rc;
  if (!h->d_LH.p || !out) return fail(TTB_EINVAL, "ttb_enqueue_fetch_site_lh: nothing computed yet");
  CK(cudaMemcpyAsync(out, h->d_LH.p, h->Lp * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

int ttb_enqueue_fetch_all_seq_idx(ttb_handle h, uint8_t* out) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_states(h)) return rc;
  if (!out) return fail(TTB_EINVAL, "ttb_enqueue_fetch_all_seq_idx: null output");
  if (int rc = h->d_bstage.alloc(std::max((size_t)h->n_tips, (size_t)h->n_int) * (size_t)h->Lp)) return rc;
  pitch_bytes_kernel<<<h->n_sm * 8, 256, 0, h->stream>>>(h->d_idx.p, h->ld, h->d_bstage.p, h->Lp, h->Lp, h->n_int, 0);
  h->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->d_bstage.p, (size_t)h->n_int * h->Lp, cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

int ttb_fetch_all_seq_idx(ttb_handle h, uint8_t* out) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_states(h)) return rc;
  if (!out) return fail(TTB_EINVAL, "ttb_fetch_all_seq_idx: null output");
  if (int rc = h->d_bstage.alloc(std::max((size_t)h->n_tips, (size_t)h->n_int) * (size_t)h->Lp)) return rc;
  pitch_bytes_kernel<<<h->n_sm * 8, 256, 0, h->stream>>>(h->d_idx.p, h->ld, h->d_bstage.p, h->Lp, h->Lp, h->n_int, 0);
  h->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->d_bstage.p, (size_t)h->n_int * h->Lp, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int ttb_profile_marginal(ttb_handle h, int32_t flags, double* ms, int32_t* launches) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, false)) return rc;
  if (h->f32 && h->have_masks) return fail(TTB_EUNSUPPORTED, "float message storage is not available together with per-branch masks");
  if (!ms || !launches) return fail(TTB_EINVAL, "ttb_profile_marginal: null output");
  const bool lh_only = flags & TTB_LH_ONLY;
  const bool tips = (flags & TTB_RECONSTRUCT_TIPS) && !lh_only;
  const bool keep_prev = (flags & TTB_KEEP_PREV_STATES) && !lh_only;
  flags = (lh_only ? TTB_LH_ONLY : 0) | (tips ? TTB_RECONSTRUCT_TIPS : 0);
  const bool had_P = h->d_P.p != nullptr;
  if (int rc = ensure_state(h, tips)) return rc;
  if (!had_P) h->drop_graphs();
  if (!lh_only)
    if (int rc = ensure_preorder_state(h, tips)) return rc;
  if (int rc = update_ss_interp(h)) return rc;
  if (keep_prev) {   // the states this pass is about to overwrite: what ttb_sample_states counts changes against
    if (int rc = h->d_idx_prev.alloc(h->d_idx.n)) return rc;
    CK(cudaMemcpyAsync(h->d_idx_prev.p, h->d_idx.p, h->d_idx.bytes(), cudaMemcpyDeviceToDevice, h->stream));
    if (tips) {
      if (int rc = h->d_idxtip_prev.alloc(h->d_idxtip.n)) return rc;
      CK(cudaMemcpyAsync(h->d_idxtip_prev.p, h->d_idxtip.p, h->d_idxtip.bytes(), cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  if (!lh_only) {
    h->have_prev = keep_prev;
    h->have_prev_tips = keep_prev && tips;
  }
  const char* trace_path = getenv("TTB_TRACE");   // measurement only: block timelines of the traced level kernels -> raw file
  const size_t trace_words = 16 + (size_t)TTB_TRACE_SLOTS * TTB_TRACE_SLOT_WORDS;
  if (trace_path) {
    if (int rc = h->d_trace.alloc(trace_words)) return rc;
    CK(cudaMemsetAsync(h->d_trace.p, 0, trace_words * 8, h->stream));
    const unsigned long long filt = getenv("TTB_TRACE_GRID") ? strtoull(getenv("TTB_TRACE_GRID"), nullptr, 10) : 0ull;   // only launches of this grid size
    CK(cudaMemcpyAsync(h->d_trace.p + 1, &filt, 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  cudaEvent_t ev[6];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  int nk = 0, pk[4] = {0, 0, 0, 0};
  int rc = enqueue_pass(h, flags, lh_only ? 0 : 1, h->stream, &nk, ev, pk);
  if (rc) return rc;
  CK(cudaGetLastError());
  if (trace_path) {
    std::vector<unsigned long long> tr(trace_words);
    CK(cudaMemcpyAsync(tr.data(), h->d_trace.p, trace_words * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (FILE* f = fopen(trace_path, "wb")) { fwrite(tr.data(), 8, trace_words, f); fclose(f); }
    h->d_trace.release();
  }
  CK(cudaMemcpyAsync(h->h_results, h->d_results.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float f;
  CK(cudaEventElapsedTime(&f, ev[0], ev[1])); ms[0] = f;
  CK(cudaEventElapsedTime(&f, ev[1], ev[2])); ms[1] = f;
  CK(cudaEventElapsedTime(&f, ev[2], ev[3])); ms[2] = f;
  CK(cudaEventElapsedTime(&f, ev[4], ev[5])); ms[2] += f;
  CK(cudaEventElapsedTime(&f, ev[3], ev[4])); ms[3] = f;
  for (int i = 0; i < 4; ++i) launches[i] = pk[i];
  for (auto& e : ev) cudaEventDestroy(e);
  h->launches += nk;
  if (!lh_only) {
    h->have_pass = true;
    h->have_tip_pass = tips;
    h->have_prev = h->have_prev_tips = false;
  }
  return 0;
}

static int branch_eval(ttb_handle h, int32_t n_eval, const int32_t* nodes, const int32_t* kind, const double* t,
                       int mode, double* out) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, true)) return rc;
  if (n_eval < 0 || (n_eval && (!nodes || !out || (mode == 0 && !t)))) return fail(TTB_EINVAL, "branch evaluation: bad arguments");
  if (n_eval == 0) return 0;
  const bool root_bif = (h->child_ptr[1] - h->child_ptr[0]) == 2;
  for (int e = 0; e < n_eval; ++e) {
    if (nodes[e] <= 0 || nodes[e] >= h->n_nodes) return fail(TTB_EINVAL, "branch evaluation: node id out of range (the root has no branch)");
    if (kind && kind[e] == TTB_BRANCH_ROOT && (!root_bif || h->parent[nodes[e]] != 0))
      return fail(TTB_EINVAL, "branch evaluation: TTB_BRANCH_ROOT needs a child of a bifurcating root");
  }
  cudaStream_t s = h->stream;
  int rc;
  if ((rc = upload(h->d_enodes, nodes, (size_t)n_eval, s))) return rc;
  if (kind) {
    if ((rc = upload(h->d_ekinds, kind, (size_t)n_eval, s))) return rc;
  }
  if (mode == 0) {
    if ((rc = upload(h->d_ets, t, (size_t)n_eval, s))) return rc;
  }
  // enough blocks per branch to fill the GPU without starving long alignments
  int nb = (int)std::min<long long>(h->tiles(), std::max<long long>(1, ((long long)h->n_sm * 16 + n_eval - 1) / n_eval));
  nb = std::min(nb, 64);
  if ((rc = h->d_partial.alloc((size_t)n_eval * nb))) return rc;
  if ((rc = h->d_eout.alloc((size_t)n_eval))) return rc;
  ttb_qops(h->q)->branch_eval(h->dev(), n_eval, nb, h->d_enodes.p, kind ? h->d_ekinds.p : nullptr, h->d_ets.p, mode,
                              h->d_partial.p, h->d_eout.p, s);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->d_eout.p, (size_t)n_eval * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

int ttb_branch_objective(ttb_handle h, int32_t n_eval, const int32_t* nodes, const int32_t* kind, const double* t, double* f) {
  return branch_eval(h, n_eval, nodes, kind, t, 0, f);
}

int ttb_branch_hamming(ttb_handle h, int32_t n_eval, const int32_t* nodes, const int32_t* kind, double* num, double* den) {
  if (int rc = branch_eval(h, n_eval, nodes, kind, nullptr, 1, num)) return rc;
  if (den) {
    double s = 0.0;
    for (double m : h->h_mult) s += m;
    *den = s;
  }
  return 0;
}

static TtbBrent brent_view(ttb_handle h) {
  TtbBrent B;
  double* d = h->d_brent.p;
  const size_t n = (size_t)h->n_brent;
  double** slots[] = {&B.xa, &B.xb, &B.xc, &B.fa, &B.fb, &B.fc, &B.x, &B.w, &B.v, &B.fx, &B.fw, &B.fv, &B.a, &B.b, &B.deltax, &B.rat, &B.u, &B.ts};
  for (size_t k = 0; k < sizeof(slots) / sizeof(slots[0]); ++k) *slots[k] = d + k * n;
  B.f = d + 18 * n;
  B.nit = h->d_brent_i.p;
  B.nfev = h->d_brent_i.p + n;
  B.flags = h->d_brent_i.p + 2 * n;
  B.active = h->d_brent_active.p;
  return B;
}

int ttb_brent_begin(ttb_handle h, int32_t n, const int32_t* nodes, const int32_t* kind, const double* xa, const double* xb,
                    const double* xc, double tol, int32_t maxiter) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, true)) return rc;
  if (n <= 0 || !nodes || !xa || !xb || !xc || !(tol > 0.0) || maxiter <= 0) return fail(TTB_EINVAL, "ttb_brent_begin: bad arguments");
  const bool root_bif = (h->child_ptr[1] - h->child_ptr[0]) == 2;
  for (int e = 0; e < n; ++e) {
    if (nodes[e] <= 0 || nodes[e] >= h->n_nodes) return fail(TTB_EINVAL, "ttb_brent_begin: node id out of range (the root has no branch)");
    if (kind && kind[e] == TTB_BRANCH_ROOT && (!root_bif || h->parent[nodes[e]] != 0))
      return fail(TTB_EINVAL, "ttb_brent_begin: TTB_BRANCH_ROOT needs a child of a bifurcating root");
  }
  cudaStream_t s = h->stream;
  int rc;
  h->n_brent = n;
  if ((rc = h->d_brent.alloc((size_t)19 * n))) return rc;
  if ((rc = h->d_brent_i.alloc((size_t)2 * n + 2))) return rc;
  if ((rc = h->d_brent_active.alloc((size_t)n))) return rc;
  if (!h->h_brent_flags) CK(cudaMallocHost(&h->h_brent_flags, 2 * sizeof(int)));
  if ((rc = upload(h->d_enodes, nodes, (size_t)n, s))) return rc;
  if (kind) {
    if ((rc = upload(h->d_ekinds, kind, (size_t)n, s))) return rc;
  } else {
    h->d_ekinds.release();
  }
  TtbBrent B = brent_view(h);
  CK(cudaMemcpyAsync(B.xa, xa, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(B.xb, xb, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(B.xc, xc, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(h->d_brent_i.p, 0, h->d_brent_i.bytes(), s));
  CK(cudaMemsetAsync(h->d_brent_active.p, 1, (size_t)n, s));
  CK(cudaStreamSynchronize(s));      // pageable sources
  h->brent_tol = tol;
  h->brent_maxiter = maxiter;
  h->brent_stage = 0;
  int nb = (int)std::min<long long>(h->tiles(), std::max<long long>(1, ((long long)h->n_sm * 16 + n - 1) / n));
  h->brent_nb = std::min(nb, 64);
  if ((rc = h->d_partial.alloc((size_t)n * h->brent_nb))) return rc;
  ttb_brent_step(B, n, -1, tol, maxiter, s);      // trial points = the first bracket point
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

int ttb_brent_eval(ttb_handle h) {
  if (int rc = use_device(h)) return rc;
  if (!h->n_brent || !h->d_brent.p) return fail(TTB_EINVAL, "ttb_brent_eval: call ttb_brent_begin first");
  if (int rc = check_ready(h, true)) return rc;
  TtbBrent B = brent_view(h);
  ttb_qops(h->q)->branch_eval(h->dev(), h->n_brent, h->brent_nb, h->d_enodes.p, h->d_ekinds.p, B.ts, 0, h->d_partial.p,
                              const_cast<double*>(B.f), h->stream);
  h->launches += 2;
  CK(cudaGetLastError());
  return 0;
}

int ttb_brent_f_device_ptr(ttb_handle h, void** dptr, int32_t* n) {
  if (int rc = use_device(h)) return rc;
  if (!h->n_brent || !h->d_brent.p || !dptr) return fail(TTB_EINVAL, "ttb_brent_f_device_ptr: call ttb_brent_begin first");
  *dptr = const_cast<double*>(brent_view(h).f);
  if (n) *n = h->n_brent;
  return 0;
}

int ttb_brent_update(ttb_handle h, int32_t sync, int32_t* n_active) {
  if (int rc = use_device(h)) return rc;
  if (!h->n_brent || !h->d_brent.p) return fail(TTB_EINVAL, "ttb_brent_update: call ttb_brent_begin first");
  TtbBrent B = brent_view(h);
  cudaStream_t s = h->stream;
  if (h->brent_stage >= 2) CK(cudaMemsetAsync(B.flags, 0, sizeof(int), s));    // the active count of this step (the error flag stays)
  ttb_brent_step(B, h->n_brent, h->brent_stage, h->brent_tol, h->brent_maxiter, s);
  h->launches += 1;
  CK(cudaGetLastError());
  h->brent_stage += 1;
  if (sync) {
    CK(cudaMemcpyAsync(h->h_brent_flags, B.flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h->brent_stage >= 3 && h->h_brent_flags[1]) {
      return fail(TTB_EINVAL, (h->h_brent_flags[1] & 1)
                                  ? "Bracketing values (xa, xb, xc) do not fulfill this requirement: (xa < xb) and (xb < xc)"
                                  : "Bracketing values (xa, xb, xc) do not fulfill this requirement: (f(xb) < f(xa)) and (f(xb) < f(xc))");
    }
    if (n_active) *n_active = h->brent_stage >= 3 ? h->h_brent_flags[0] : h->n_brent;
  } else if (n_active) {
    *n_active = -1;
  }
  return 0;
}

int ttb_brent_result(ttb_handle h, double* x, double* fun, int32_t* nit, int32_t* nfev) {
  if (int rc = use_device(h)) return rc;
  if (!h->n_brent || !h->d_brent.p || h->brent_stage < 3) return fail(TTB_EINVAL, "ttb_brent_result: no finished minimisation");
  TtbBrent B = brent_view(h);
  const size_t n = (size_t)h->n_brent;
  cudaStream_t s = h->stream;
  if (x) CK(cudaMemcpyAsync(x, B.x, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (fun) CK(cudaMemcpyAsync(fun, B.fx, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (nit) CK(cudaMemcpyAsync(nit, B.nit, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (nfev) CK(cudaMemcpyAsync(nfev, B.nfev, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

int ttb_branch_state_pairs(ttb_handle h, int32_t n, const int32_t* nodes, int32_t tip_states, int32_t width, double* counts,
                           int32_t* first) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_states(h)) return rc;
  if (n <= 0 || !nodes || !counts || !first) return fail(TTB_EINVAL, "ttb_branch_state_pairs: bad arguments");
  const int need = tip_states ? h->q : std::max(h->q, h->n_codes);
  if (width < need) return fail(TTB_EINVAL, "ttb_branch_state_pairs: width must be >= max(n_states, n_codes)");
  bool any_tip = false;
  for (int k = 0; k < n; ++k) {
    if (nodes[k] <= 0 || nodes[k] >= h->n_nodes) return fail(TTB_EINVAL, "ttb_branch_state_pairs: node index out of range (the root has no branch)");
    any_tip |= h->tip_row[nodes[k]] >= 0;
  }
  if (tip_states && any_tip && !(h->have_tip_pass || h->have_joint_tips))
    return fail(TTB_EINVAL, "ttb_branch_state_pairs: tip states requested but the last pass did not reconstruct tips");
  const size_t bins = (size_t)h->q * width;
  const size_t smem = bins * (sizeof(double) + sizeof(int));
  if (smem > 200 * 1024) return fail(TTB_EUNSUPPORTED, "ttb_branch_state_pairs: table too large for shared memory");
  int rc;
  if ((rc = h->d_enodes.alloc((size_t)n))) return rc;
  if ((rc = h->d_pair_counts.alloc((size_t)n * bins))) return rc;
  if ((rc = h->d_pair_first.alloc((size_t)n * bins))) return rc;
  CK(cudaMemcpyAsync(h->d_enodes.p, nodes, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(pair_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_counts_kernel<<<n, 256, smem, h->stream>>>(h->dev(), h->d_enodes.p, tip_states ? 1 : 0, width, h->d_pair_counts.p, h->d_pair_first.p);
  h->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(counts, h->d_pair_counts.p, (size_t)n * bins * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(first, h->d_pair_first.p, (size_t)n * bins * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int ttb_seqgen(ttb_handle h, uint64_t seed, const uint8_t* root_idx, const double* uniforms, const uint8_t* state2code,
               uint8_t* states_out) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, false)) return rc;
  if (!state2code) return fail(TTB_EINVAL, "ttb_seqgen: state2code is null");
  for (int i = 0; i < h->q; ++i)
    if (state2code[i] >= h->n_codes) return fail(TTB_EINVAL, "ttb_seqgen: state2code entry out of range of the code table");
  if (root_idx)
    for (long long a = 0; a < h->Lp; ++a)
      if (root_idx[a] >= h->q) return fail(TTB_EINVAL, "ttb_seqgen: root state out of range");
  if (int rc = update_ss_interp(h)) return rc;
  const size_t q = h->q, Lp = (size_t)h->Lp, ld = (size_t)h->ld;
  int rc;
  if (!h->site_specific) {
    const size_t pq = (q * q + 1) / 2 * 2;
    const bool had_P = h->d_P.p != nullptr;
    if ((rc = h->d_P.alloc((size_t)h->n_nodes * pq))) return rc;
    if (!had_P) h->drop_graphs();
  }
  if ((rc = h->d_sg_states.alloc((size_t)h->n_nodes * ld))) return rc;
  if ((rc = h->d_bstage.alloc(std::max(std::max((size_t)h->n_tips, (size_t)h->n_int) * Lp, Lp + 256)))) return rc;
  cudaStream_t s = h->stream;
  const uint8_t* d_root = nullptr;
  CK(cudaMemcpyAsync(h->d_bstage.p, state2code, q, cudaMemcpyHostToDevice, s));
  if (root_idx) {
    CK(cudaMemcpyAsync(h->d_bstage.p + 256, root_idx, Lp, cudaMemcpyHostToDevice, s));
    d_root = h->d_bstage.p + 256;
  }
  const double* d_uni = nullptr;
  if (uniforms) {
    if ((rc = h->d_sg_uniforms.alloc((size_t)h->n_nodes * Lp))) return rc;
    CK(cudaMemcpyAsync(h->d_sg_uniforms.p, uniforms, (size_t)h->n_nodes * Lp * sizeof(double), cudaMemcpyHostToDevice, s));
    d_uni = h->d_sg_uniforms.p;
  }
  const int nk = ttb_qops(h->q)->seqgen(h->dev(), h->tiles(), (unsigned long long)seed, d_root, d_uni, h->d_sg_states.p, s);
  seqgen_tip_codes_kernel<<<dim3(h->tiles(), std::min(h->n_tips, h->n_sm * 4)), TTB_BLOCK, 0, s>>>(h->dev(), h->d_tip_nodes.p, h->d_sg_states.p,
                                                                                            h->d_bstage.p, h->d_codes.p);
  h->launches += nk + 1;
  CK(cudaGetLastError());
  // the alignment changed: no reconstruction is valid any more, previous states are meaningless
  if (h->d_idx.p) CK(cudaMemsetAsync(h->d_idx.p, 0xff, h->d_idx.bytes(), s));
  if (h->d_idxtip.p) CK(cudaMemsetAsync(h->d_idxtip.p, 0xff, h->d_idxtip.bytes(), s));
  h->have_pass = h->have_tip_pass = false;
  h->have_joint = h->have_joint_tips = false;
  if (states_out)
    CK(cudaMemcpy2DAsync(states_out, Lp, h->d_sg_states.p, ld, Lp, (size_t)h->n_nodes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  h->d_sg_uniforms.release();
  return 0;
}

int ttb_mutation_counts(ttb_handle h, double* n_ij, double* T_i) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, true)) return rc;
  if (!n_ij || !T_i) return fail(TTB_EINVAL, "ttb_mutation_counts: null output");
  if (h->site_specific) return fail(TTB_EUNSUPPORTED, "ttb_mutation_counts: per-site statistics of site-specific models are not implemented");
  const int q = h->q, width = q * q + q;
  const int tiles = h->tiles();
  int chunks = std::max(1, std::min(h->n_nodes - 1, (h->n_sm * 8 + tiles - 1) / tiles));
  const int chunk = (h->n_nodes - 1 + chunks - 1) / chunks;
  chunks = (h->n_nodes - 1 + chunk - 1) / chunk;
  int rc;
  if ((rc = h->d_partial.alloc((size_t)chunks * tiles * width))) return rc;
  if ((rc = h->d_eout.alloc((size_t)width))) return rc;
  ttb_qops(h->q)->counts(h->dev(), tiles, chunks, chunk, h->d_partial.p, h->d_eout.p, h->stream);
  h->launches += 2;
  CK(cudaGetLastError());
  std::vector<double> tmp(width);
  CK(cudaMemcpyAsync(tmp.data(), h->d_eout.p, width * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(n_ij, tmp.data(), sizeof(double) * q * q);
  memcpy(T_i, tmp.data() + q * q, sizeof(double) * q);
  return 0;
}

int ttb_mutation_counts_per_site(ttb_handle h, double* n_ija, double* T_ia) {
  if (int rc = use_device(h)) return rc;
  if (int rc = check_ready(h, true)) return rc;
  if (!n_ija || !T_ia) return fail(TTB_EINVAL, "ttb_mutation_counts_per_site: null output");
  const int q = h->q, width = q * q + q;
  const int tiles = h->tiles();
  const size_t ld = (size_t)h->ld, Lp = (size_t)h->Lp;
  // enough branch chunks to fill the GPU, bounded by 1 GB of partial sums
  int chunks = std::max(1, std::min(h->n_nodes - 1, (h->n_sm * 8 + tiles - 1) / tiles));
  chunks = (int)std::max<size_t>(1, std::min<size_t>((size_t)chunks, ((size_t)1 << 27) / ((size_t)width * ld)));
  const int chunk = (h->n_nodes - 1 + chunks - 1) / chunks;
  chunks = (h->n_nodes - 1 + chunk - 1) / chunk;
  int rc;
  if ((rc = h->d_partial.alloc((size_t)chunks * width * ld))) return rc;
  if ((rc = h->d_stage.alloc((size_t)width * ld))) return rc;
  ttb_qops(h->q)->site_counts(h->dev(), tiles, chunks, chunk, h->d_partial.p, h->d_stage.p, h->stream);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpy2DAsync(n_ija, Lp * sizeof(double), h->d_stage.p, ld * sizeof(double), Lp * sizeof(double), (size_t)q * q,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpy2DAsync(T_ia, Lp * sizeof(double), h->d_stage.p + (size_t)q * q * ld, ld * sizeof(double), Lp * sizeof(double), (size_t)q,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->d_partial.release();
  return 0;
}

int ttb_device_bytes(ttb_handle h, int64_t* bytes) {
  if (!h || !bytes) return fail(TTB_EINVAL, "null argument");
  size_t b = h->d_codes.bytes() + h->d_code_prof.bytes() + h->d_mult.bytes() + h->d_TU.bytes() + h->d_t.bytes() +
             h->d_P.bytes() + h->d_Pf.bytes() + h->d_S.bytes() + h->d_F.bytes() + h->d_M.bytes() + h->d_Mtip.bytes() + h->d_LH.bytes() +
             h->d_idx.bytes() + h->d_idxtip.bytes() + h->d_stage.bytes() + h->d_partial.bytes() + h->d_parent.bytes() * 5;
  *bytes = (int64_t)b;
  return 0;
}

int ttb_launch_count(ttb_handle h, int64_t* count) {
  if (!h || !count) return fail(TTB_EINVAL, "null argument");
  *count = h->launches;
  return 0;
}

}  // extern "C"
