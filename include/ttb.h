/*
 * ttb.h -- C-ABI of the B200 marginal ancestral-reconstruction engine (libttb.so).
 *
 * This is the drop-in boundary for ONE hot path of neherlab/treetime:
 * TreeAnc._ml_anc_marginal (treetime/treeanc.py:762-932) and the branch-length /
 * likelihood surface that consumes its messages (treeanc.py:1085-1146,1272-1360,
 * gtr.py:816-963).  The reference has no FFI (it is pure Python/numpy); every entry
 * point below names the reference code it replaces.  INTEGRATION.md shows the ctypes
 * binding a TreeTime maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative TTB_E* code; ttb_last_error()
 *    returns a thread-local description.  No C++ exception crosses the boundary.
 *  - the caller owns every host buffer; the library owns device memory behind the handle.
 *  - one handle = one GPU = one shard of the pattern axis.  Calls on a handle are
 *    stream-ordered and not thread-safe.  Only ttb_fetch_* / ttb_results / ttb_sync block.
 *  - nodes are numbered in the reference's preorder (tree.find_clades()), root = 0;
 *    children are listed in `node.clades` order.
 *  - host matrices handed in/out are row-major (L', q) like the reference's numpy arrays;
 *    on the device messages are state-planar [node][state][pattern].
 */
#ifndef TTB_H
#define TTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ttb_engine* ttb_handle;

enum {
  TTB_OK = 0,
  TTB_EINVAL = -1,   /* bad argument / call order */
  TTB_ECUDA = -2,    /* CUDA runtime error (message has the CUDA string) */
  TTB_ENOMEM = -3,   /* device allocation failed */
  TTB_EUNSUPPORTED = -4 /* e.g. n_states without a compiled kernel */
};

/* which per-node array ttb_fetch_node returns */
enum {
  TTB_SUBTREE = 0,  /* node.marginal_subtree_LH   (treeanc.py:877)  */
  TTB_OUTGROUP = 1, /* node.marginal_outgroup_LH  (treeanc.py:895-899) */
  TTB_PROFILE = 2,  /* node.marginal_profile      (treeanc.py:822-824,910-912) */
  TTB_JOINT_ROOT_LX = 3 /* root.joint_Lx after ttb_joint (treeanc.py:1003-1004); node must be 0 */
};

/* flags of ttb_marginal */
enum {
  TTB_RECONSTRUCT_TIPS = 1, /* reconstruct_tip_states=True (treeanc.py:900-903) */
  TTB_LH_ONLY = 2,          /* postorder + root only: the cost function of optimize_gtr_rate (treeanc.py:1685-1689) */
  TTB_JOINT_NO_TRACE = 4,   /* ttb_joint: stop after the root (the caller samples the root, then ttb_joint_retrace) */
  TTB_KEEP_PREV_STATES = 32 /* ttb_marginal: keep a copy of the states this pass overwrites (needed by ttb_sample_states) */
};

/* kinds of branch evaluated by ttb_branch_objective / ttb_branch_hamming */
enum {
  TTB_BRANCH = 0,      /* (pp, pc) = (outgroup_LH(n), subtree_LH(n))            treeanc.py:1122-1146 */
  TTB_BRANCH_ROOT = 1  /* merged branch across a bifurcating root, n = first root child:
                          pc = subtree(n1), pp = normalize(subtree(n2) * Pi)     treeanc.py:1317-1326 */
};

const char* ttb_last_error(void);
int ttb_version(void);
/* 1 if kernels for this alphabet size are compiled in (2..8 and 20..22) */
int ttb_supports_n_states(int n_states);

/* Create an engine on CUDA device `device` for an alphabet of n_states (gtr.n_states). */
int ttb_create(ttb_handle* out, int device, int n_states);
int ttb_destroy(ttb_handle h);
/* Run on an externally owned stream (e.g. torch's current stream); NULL = engine's own stream. */
int ttb_set_stream(ttb_handle h, void* cuda_stream);

/* Storage type of the per-node message arrays (marginal_subtree_LH, marginal_profile) in device memory.  The reference
 * keeps them as float64 numpy arrays (treeanc.py:877,910-912) and that is the default.  TTB_STORAGE_F32 stores them as
 * float while every arithmetic operation stays in double: the level kernels move half the bytes.  Opt-in, with a
 * measured tolerance (total log-LH ~1e-8 relative instead of 1e-16, profiles ~1e-7; sequences can
 * differ where the two largest profile entries are closer than ~1e-7).  Alphabets of up to 8 states; not with
 * per-branch masks or ttb_joint.  Changing the type invalidates the current reconstruction. */
enum { TTB_STORAGE_F64 = 0, TTB_STORAGE_F32 = 1 };
int ttb_set_message_storage(ttb_handle h, int32_t storage);

/* Tree topology (replaces the Bio.Phylo walk of treeanc.py:857,887).  parent[root] = -1.
 * tip_row[n] = row of the tip-code matrix for terminal nodes, -1 for internal nodes.
 * The level schedules (by height for the postorder, by depth for the preorder) are
 * built inside. */
int ttb_set_tree(ttb_handle h, int32_t n_nodes, const int32_t* parent, const int32_t* child_ptr,
                 const int32_t* child_idx, const int32_t* tip_row);

/* Compressed alignment shard (replaces seq2prof on the leaves, treeanc.py:846-853):
 * tip_codes[n_tips][n_patterns] uint8 indices into code_profiles[n_codes][n_states]
 * (the values of gtr.profile_map), multiplicity[n_patterns] = data.multiplicity(). */
int ttb_set_patterns(ttb_handle h, int64_t n_patterns, const uint8_t* tip_codes, int32_t n_codes,
                     const double* code_profiles, const double* multiplicity);

/* Sparse variant of ttb_set_patterns (the analogue of TreeTime's dict-of-differences / VCF alignments,
 * sequence_data.py:363-383): every tip row equals ref_codes[n_patterns] except at the n_entries listed
 * (tip_row, pattern, code) positions.  The dense code matrix is built on the device. */
int ttb_set_patterns_sparse(ttb_handle h, int64_t n_patterns, const uint8_t* ref_codes, int64_t n_entries,
                            const int32_t* entry_row, const int32_t* entry_pos, const uint8_t* entry_code,
                            int32_t n_codes, const double* code_profiles, const double* multiplicity);

/* N3 -- pattern compression on the device (SequenceData.make_compressed_alignment, sequence_data.py:325-464).
 * Step 1: upload the raw alignment aln[n_seq][L] (ASCII, already upper-cased), optionally turn leading /
 * trailing `gap` characters into `fill` (seq2array fill_overhangs, seq_utils.py:196-202; fill_overhangs = 0
 * skips it) and return per column the extrema over the characters != `ambiguous` plus an all-ambiguous flag:
 * a column is constant iff lo == hi (or all ambiguous).  The alignment stays resident on the device. */
int ttb_alignment_stats(ttb_handle h, int64_t n_seq, int64_t L, const uint8_t* aln, int32_t fill_overhangs,
                        int32_t gap, int32_t fill, int32_t ambiguous, uint8_t* lo, uint8_t* hi, uint8_t* all_amb);
/* Step 2 (after the host numbered the patterns): pattern p is the column first_pos[p] of the resident
 * alignment, or -- if const_letter[p] != 0 -- a constant column of that letter; tip t reads alignment row
 * tip_seq_row[t] (-1 = no sequence = missing_code); lut[256] maps characters to codes (255 = unknown
 * character -> TTB_EINVAL).  Equivalent to ttb_set_patterns with the gathered code matrix. */
int ttb_set_patterns_from_alignment(ttb_handle h, int64_t n_patterns, const int64_t* first_pos, const uint8_t* const_letter,
                                    const int32_t* tip_seq_row, const uint8_t* lut, int32_t missing_code, int32_t n_codes,
                                    const double* code_profiles, const double* multiplicity);

/* Single-site GTR eigen-system (gtr.py:612-629): eigvals[q], v[q][q], v_inv[q][q], Pi[q], mu.
 * gap_index = gtr.gap_index or -1 (used by the branch objective, gtr.py:954-959). */
int ttb_set_gtr(ttb_handle h, const double* eigvals, const double* v, const double* v_inv,
                const double* Pi, double mu, int32_t gap_index);

/* Site-specific GTR (gtr_site_specific.py:312-371), arrays in the reference's layout for
 * this shard's patterns: eigvals[q][L'], v[q][q][L'], v_inv[q][q][L'], Pi[q][L'], mu[L'];
 * t_grid[n_grid] is the interpolation grid (:336-344); approximate != 0 selects the
 * linear-in-t interpolated expQt for t*rate_scale < 10 (:367-371). */
int ttb_set_gtr_site_specific(ttb_handle h, const double* eigvals, const double* v, const double* v_inv,
                              const double* Pi, const double* mu, const double* t_grid, int32_t n_grid,
                              double rate_scale, int32_t approximate, int32_t gap_index);

/* t[n_nodes]: branch lengths as the GTR sees them, i.e. after _branch_length_to_gtr
 * (treeanc.py:752-760).  t[root] is ignored. */
int ttb_set_branch_lengths(ttb_handle h, const double* t);

/* Per-branch masks of the ARG mode (arg.py:128-133; node.mask, treeanc.py:454,489-490): masks[n_masks][n_patterns]
 * with entries 0 / 1, node_mask[n_nodes] = row of `masks` used by the branch above that node or -1 (node.mask is None).
 * A masked (branch, pattern) passes no information: the child's up-message is dropped (treeanc.py:862-872), the
 * child's profile is its subtree profile (:909-917), and the pattern's multiplicity is zero in that branch's
 * likelihood and substitution statistics (data.multiplicity(mask=node.mask), :1294,1326-1333,1564-1572).
 * n_masks = 0 removes all masks.  A new tree or alignment also removes them.  Not available for ttb_joint. */
int ttb_set_branch_masks(ttb_handle h, int32_t n_masks, const uint8_t* masks, const int32_t* node_mask);

/* Enqueue one marginal reconstruction: batched expQt, level-ordered postorder, root,
 * level-ordered preorder (treeanc.py:762-812), one CUDA graph launch.  Asynchronous. */
int ttb_marginal(ttb_handle h, int32_t flags);
/* N2 -- joint (max-product) ML reconstruction, TreeAnc._ml_anc_joint (treeanc.py:934-1080): log-space
 * postorder with back-pointers, root state, back-trace; flags: TTB_RECONSTRUCT_TIPS.  Results through
 * ttb_results (total = tree.sequence_joint_LH, n_diff), ttb_fetch_site_lh (tree.sequence_LH of the joint
 * path, :1021) and the ttb_fetch_*seq_idx / ttb_fetch_mutations calls.  It overwrites the marginal
 * messages: marginal accessors need a new ttb_marginal afterwards.  Not available for site-specific models. */
int ttb_joint(ttb_handle h, int32_t flags);
/* Back-trace from caller-chosen root states root_idx[n_patterns] (sample_from_profile='root',
 * treeanc.py:1008-1023); needs a preceding ttb_joint (usually with TTB_JOINT_NO_TRACE). */
int ttb_joint_retrace(ttb_handle h, const uint8_t* root_idx, int32_t flags);

/* sample_from_profile=True (treeanc.py:786-798,919-923): replace the argmax states of the n listed nodes by states
 * drawn from their marginal profiles exactly like prof2seq (seq_utils.py:266-269): state = first i with
 * cumsum(profile)[i] >= u, state 0 if there is none.  uniforms[n][n_patterns] holds the caller's draws, row k for
 * nodes[k] -- for the reference's sequences pass rng.random(L') per reconstructed non-root node in preorder, after
 * the root's own draw.  The profiles do not depend on drawn states, so this follows a ttb_marginal run with
 * TTB_KEEP_PREV_STATES; n_diff / n_diff_tips receive the number of (internal / tip node, pattern) states that
 * differ from the states before that pass (treeanc.py:925-926) and replace the pass's own counts.  Tips need
 * TTB_RECONSTRUCT_TIPS.  Synchronous. */
int ttb_sample_states(ttb_handle h, int32_t n, const int32_t* nodes, const double* uniforms, int64_t* n_diff,
                      int64_t* n_diff_tips);

/* Wait for the last ttb_marginal and return this shard's partial results:
 * total_lh = sum_a LH_a * multiplicity_a (treeanc.py:828), n_diff = number of (node, pattern)
 * state indices that changed w.r.t. the previous reconstruction (treeanc.py:925-926). */
int ttb_results(ttb_handle h, double* total_lh, int64_t* n_diff);
/* The part of n_diff that belongs to terminal nodes (TTB_RECONSTRUCT_TIPS passes). */
int ttb_results_tips(ttb_handle h, int64_t* n_diff_tips);
/* Device address of the double[2] {total_lh, n_diff} written by the last pass (for NCCL allreduce). */
int ttb_results_device_ptr(ttb_handle h, void** dptr);
int ttb_sync(ttb_handle h);

/* tree.sequence_LH (treeanc.py:825-827): out[n_patterns]. */
int ttb_fetch_site_lh(ttb_handle h, double* out);
/* One per-node array, row-major out[n_patterns][n_states]. */
int ttb_fetch_node(ttb_handle h, int32_t node, int32_t which, double* out);
/* argmax state indices (alphabet[idx] = node._cseq, seq_utils.py:271) for `n` nodes:
 * out[n][n_patterns].  Tips only after TTB_RECONSTRUCT_TIPS. */
int ttb_fetch_seq_idx(ttb_handle h, int32_t n, const int32_t* nodes, uint8_t* out);

/* All internal nodes at once: out[n_internal][n_patterns], rows in node (preorder) order. */
int ttb_fetch_all_seq_idx(ttb_handle h, uint8_t* out);

/* Sparse form of the reconstructed sequences: root_idx[n_patterns] plus every (node, pattern, state) where
 * an internal node's state differs from its parent's (what `node.mutations` lists, treeanc.py:27-42, on
 * compressed patterns).  The first max_n entries are written, ordered by (node, pattern) -- the device lays them out
 * in that order (count per node, scan, ordered write; no sort anywhere); *n receives the total count, so a caller that
 * passed too small a buffer can retry.  Synchronous. */
int ttb_fetch_mutations(ttb_handle h, uint8_t* root_idx, int32_t max_n, int32_t* node, int32_t* pos, uint8_t* state,
                        int64_t* n);

/* N2: sufficient statistics of the joint branch-length optimisation (TreeAnc.add_branch_state,
 * treeanc.py:1148-1163; GTR.state_pair, gtr.py:631-705) for n branches given by their child nodes:
 * counts[b][i][c] = sum of multiplicities of this shard's patterns where the parent's reconstructed state is i
 * and the child shows c; c is the child's reconstructed state index for internal nodes (and for tips when
 * tip_states != 0, which needs a TTB_RECONSTRUCT_TIPS pass), otherwise the tip's alignment code, so that the
 * caller can treat ambiguous characters as the reference does.  first[b][i][c] = first pattern with that pair
 * (0x7fffffff if none): the reference lists pairs of large alphabets in order of first occurrence.
 * Rows are `width` wide, width >= max(n_states, n_codes) (n_states with tip_states).  Synchronous. */
int ttb_branch_state_pairs(ttb_handle h, int32_t n, const int32_t* nodes, int32_t tip_states, int32_t width, double* counts,
                           int32_t* first);

/* N4: evolve sequences down the tree on the device (SeqGen.evolve, seqgen.py:38-67): root ~ Pi unless
 * root_idx[n_patterns] is given, every child state drawn from column (parent state) of exp(Q t_child) with
 * argmax(cumsum(p) > u) (seqgen.py:19-36).  u comes from `uniforms` ([n_nodes][n_patterns], node order of
 * ttb_set_tree; pass the reference's draws to reproduce its sequences exactly) or, if null, from a
 * counter-based Philox4x32-10 stream keyed by `seed` (counter = site, node).  Needs tree, patterns (only their
 * number and the code table matter; multiplicities should be 1), model and branch lengths.  The tips' states
 * become the engine's alignment (code = state2code[state]); states_out (optional, [n_nodes][n_patterns])
 * receives every node's state index.  Synchronous. */
int ttb_seqgen(ttb_handle h, uint64_t seed, const uint8_t* root_idx, const double* uniforms, const uint8_t* state2code,
               uint8_t* states_out);

/* Stream-ordered variants without a host sync: the data is valid after ttb_sync / ttb_results.
 * `out` should be page-locked memory (otherwise the driver stages and the call blocks).  With
 * page-locked INPUT buffers ttb_set_patterns is asynchronous too (when sizes are unchanged): the
 * caller must then leave them untouched until the next synchronising call. */
int ttb_enqueue_fetch_site_lh(ttb_handle h, double* out);
int ttb_enqueue_fetch_all_seq_idx(ttb_handle h, uint8_t* out);

/* Same work as ttb_marginal but launched kernel by kernel with CUDA events between the
 * phases (no graph): ms[4] = {expQt batch, postorder levels, root + reductions, preorder levels},
 * launches[4] = kernels per phase.  Measurement aid for bench.py's roofline. Synchronous. */
int ttb_profile_marginal(ttb_handle h, int32_t flags, double* ms, int32_t* launches);

/* Branch-length likelihood surface: f[e] = prob_t_profiles((pp,pc), multiplicity, t[e],
 * return_log=True) (gtr.py:922-963) for the branch above nodes[e] (kind[e], may be NULL = all
 * TTB_BRANCH), using the messages of the last ttb_marginal.  Partial sum over this shard. */
int ttb_branch_objective(ttb_handle h, int32_t n_eval, const int32_t* nodes, const int32_t* kind,
                         const double* t, double* f);
/* num[e] = sum_a m_a (pp_a . pc_a) for the same branches; *den = sum_a m_a
 * (hamming_distance = 1 - num/den, gtr.py:871-874). */
int ttb_branch_hamming(ttb_handle h, int32_t n_eval, const int32_t* nodes, const int32_t* kind,
                       double* num, double* den);

/* A8 on the device: GTR.optimal_t_compressed(profiles=True) (gtr.py:816-920) for n branches at once -- scipy's
 * minimize_scalar(method='brent', bracket=(xa, xb, xc), tol) on cost(s) = -prob_t_profiles(.., s^2, return_log=True)
 * + exp(s^4 / 10000) (gtr.py:876-891) -- as a lock-step state machine whose state stays on the device:
 *   ttb_brent_begin   bracket in s = sqrt(t) per branch (the reference: xa = -sqrt(MAX_BRANCH_LENGTH), xb = sqrt(hamming
 *                     distance), xc = +sqrt(MAX_BRANCH_LENGTH)), tolerance, iteration cap (scipy: 500);
 *   ttb_brent_eval    enqueue the objective of every unfinished branch at its current trial length (one launch over this
 *                     shard's patterns; partial sums when the patterns are sharded);
 *   ttb_brent_f_device_ptr   the n objective values on the device, for an all-reduce between eval and update;
 *   ttb_brent_update  enqueue the state update + next proposal; with sync != 0 it waits and returns the number of branches
 *                     still active (the first three updates consume the bracket points; an invalid bracket fails with
 *                     scipy's message);
 *   ttb_brent_result  optimum s (t = s^2), cost there, iterations and function evaluations per branch.
 * Uses the messages of the last ttb_marginal.  All calls but update(sync) / result are asynchronous. */
int ttb_brent_begin(ttb_handle h, int32_t n, const int32_t* nodes, const int32_t* kind, const double* xa, const double* xb,
                    const double* xc, double tol, int32_t maxiter);
int ttb_brent_eval(ttb_handle h);
int ttb_brent_f_device_ptr(ttb_handle h, void** dptr, int32_t* n);
int ttb_brent_update(ttb_handle h, int32_t sync, int32_t* n_active);
int ttb_brent_result(ttb_handle h, double* x, double* fun, int32_t* nit, int32_t* nfev);

/* Expected substitution statistics of infer_gtr(marginal=True) (treeanc.py:1556-1572):
 * n_ij[q][q] and T_i[q] summed over this shard's patterns and all branches. */
int ttb_mutation_counts(ttb_handle h, double* n_ij, double* T_i);

/* The same statistics per pattern, before the sum over patterns: n_ija[q][q][n_patterns] and T_ia[q][n_patterns]
 * (treeanc.py:1551-1572) -- the input of GTR_site_specific.infer (gtr_site_specific.py:207-310), i.e. of
 * infer_gtr(marginal=True, site_specific=True).  Works for single and for site-specific current models (the latter
 * with the per-pattern transition matrices, treeanc.py:1107-1108). */
int ttb_mutation_counts_per_site(ttb_handle h, double* n_ija, double* T_ia);

/* Bytes of device memory currently held by the handle. */
int ttb_device_bytes(ttb_handle h, int64_t* bytes);
/* Kernel launches issued by the library since creation (bench.py's gpu_launches). */
int ttb_launch_count(ttb_handle h, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* TTB_H */
