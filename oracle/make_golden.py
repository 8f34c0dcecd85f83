"""TEST-ONLY: generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/bioshim) in the build container.

    python oracle/make_golden.py            # writes tests/golden/

Each fixture stores the inputs (newick, alignment, model parameters) and what the
reference computed from them.  The GPU box has no reference; there the parity tests
compare the CUDA path and the oracle against these files.  Big cases store only
seeds + a checksum of the regenerated input and the reference's per-pattern output.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

refenv.activate()
from treetime_b200 import synth  # noqa: E402
from treetime_b200.flatten import flatten_treeanc  # noqa: E402

OUT = os.path.join(refenv.REPO, 'tests', 'golden')
ONLY = set(a[7:] for a in sys.argv if a.startswith('--only='))     # e.g. --only=poly70: regenerate one fixture


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gtr_params(gtr):
    d = dict(gtr_W=np.array(gtr.W), gtr_Pi=np.array(gtr.Pi), gtr_mu=np.array(gtr.mu), gtr_alphabet=np.array(gtr.alphabet),
             gtr_eigenvals=np.array(gtr.eigenvals), gtr_v=np.array(gtr.v), gtr_v_inv=np.array(gtr.v_inv))
    chars = sorted(gtr.profile_map.keys())
    d['gtr_prof_chars'] = np.array(chars)
    d['gtr_prof_table'] = np.array([gtr.profile_map[c] for c in chars], dtype=float)
    d['gtr_ambiguous'] = np.array(gtr.ambiguous if gtr.ambiguous is not None else '')
    if getattr(gtr, 'is_site_specific', False):
        d['gtr_rate_scale'] = np.array(gtr.rate_scale)
        d['gtr_approximate'] = np.array(gtr.approximate)
    return d


def run_case(name, newick, aln, gtr, store_every=1, reconstruct_tips=False, optimize=False, infer_gtr=False,
             n_bl=6, compress=True, store_inputs=True, extra=None, cseq_as_idx=False):
    if ONLY and name not in ONLY:
        return
    t0 = time.time()
    pre = gtr_params(gtr)       # profile map / ambiguous as the user built the model (before extend_profile)
    tt = refenv.reference_treeanc(newick, aln, gtr, rng_seed=1, compress=compress)
    topo, flat, g = flatten_treeanc(tt)
    N1 = tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=reconstruct_tips)
    names = [n.name for n in topo.nodes]
    out = dict(case=np.array(name), node_names=np.array(names), N_diff_first=np.array(N1),
               multiplicity=np.array(tt.data.multiplicity()), sequence_LH=np.array(tt.tree.sequence_LH),
               total_LH=np.array(tt.tree.total_sequence_LH), reconstruct_tips=np.array(reconstruct_tips),
               compress=np.array(compress), t=flat['t'])
    out.update(gtr_params(tt.gtr))
    out['gtr_prof_chars'] = pre['gtr_prof_chars']; out['gtr_prof_table'] = pre['gtr_prof_table']; out['gtr_ambiguous'] = pre['gtr_ambiguous']
    if store_inputs:
        out['newick'] = np.array(newick)
        out['aln_names'] = np.array(sorted(aln))
        out['aln_seqs'] = np.array([''.join(aln[k]) for k in sorted(aln)])
        # flat view as produced from the reference's own objects (pins flatten + tip encoding)
        out.update({'flat_' + k: v for k, v in flat.items()})
    cseq = []
    stored = []
    for i, n in enumerate(topo.nodes):
        c = n.cseq
        cseq.append(''.join(c) if c is not None else '')
        if i % store_every == 0:
            stored.append(i)
            out['subtree_%d' % i] = n.marginal_subtree_LH
            if n.up is not None:
                out['outgroup_%d' % i] = n.marginal_outgroup_LH
            if hasattr(n, 'marginal_profile'):
                out['profile_%d' % i] = n.marginal_profile
    out['stored_nodes'] = np.array(stored)
    if cseq_as_idx:
        # big cases: uint8 state indices [n_nodes, L'] (255 = no sequence) instead of strings
        lut = {c: i for i, c in enumerate(tt.gtr.alphabet)}
        M = np.full((len(cseq), len(tt.data.multiplicity())), 255, dtype=np.uint8)
        for i, n in enumerate(topo.nodes):
            if not n.is_terminal():
                M[i] = np.vectorize(lut.get)(n.cseq).astype(np.uint8)
        out['cseq_idx'] = M
    else:
        out['cseq'] = np.array(cseq)
    out['N_diff_second'] = np.array(tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=reconstruct_tips))
    bl_nodes = list(range(1, topo.n_nodes, max(1, topo.n_nodes // n_bl)))[:n_bl]
    out['bl_nodes'] = np.array(bl_nodes)
    out['bl_opt'] = np.array([tt.optimal_marginal_branch_length(topo.nodes[i]) for i in bl_nodes])
    out['bl_opt_tol1e-2'] = np.array([tt.optimal_marginal_branch_length(topo.nodes[i], tol=1e-2) for i in bl_nodes])
    # objective values of prob_t_profiles on a small t grid
    tgrid = np.array([1e-5, 1e-3, 0.02, 0.3])
    out['obj_t'] = tgrid
    out['obj'] = np.array([[tt.gtr.prob_t_profiles(tt.marginal_branch_profile(topo.nodes[i]), tt.data.multiplicity(), t,
                                                   return_log=True) for t in tgrid] for i in bl_nodes])
    if infer_gtr:
        n_ija, T_ia = None, None
        # the accumulation of treeanc.py:1556-1572
        q = tt.gtr.n_states
        L = len(tt.data.multiplicity())
        n_ija = np.zeros((q, q, L)); T_ia = np.zeros((q, L))
        for node in tt.tree.get_nonterminals():
            for c in node:
                ms = np.transpose(tt.get_branch_mutation_matrix(c, full_sequence=False), (1, 2, 0))
                T_ia += 0.5 * tt._branch_length_to_gtr(c) * ms.sum(axis=0) * tt.data.multiplicity(mask=c.mask)
                T_ia += 0.5 * tt._branch_length_to_gtr(c) * ms.sum(axis=1) * tt.data.multiplicity(mask=c.mask)
                n_ija += ms * tt.data.multiplicity(mask=c.mask)
        out['n_ij'] = n_ija.sum(axis=-1)
        out['T_i'] = T_ia.sum(axis=-1)
    if optimize:
        tt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=False, prune_short=False)
        out['opt_branch_length'] = np.array([n.branch_length for n in topo.nodes])
        out['opt_total_LH'] = np.array(tt.tree.total_sequence_LH)
        if infer_gtr:
            tt.infer_gtr(marginal=True)
            out['inferred_W'] = np.array(tt.gtr.W); out['inferred_Pi'] = np.array(tt.gtr.Pi); out['inferred_mu'] = np.array(tt.gtr.mu)
    if extra:
        out.update(extra)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-14s nodes=%5d L\'=%5d total_LH=%.6f  %.1fs  %d KB' % (name, topo.n_nodes, len(tt.data.multiplicity()),
                                                                 tt.tree.total_sequence_LH if not optimize else out['total_LH'],
                                                                 time.time() - t0, os.path.getsize(path) // 1024))


def run_joint_case(src, gtr, reconstruct_tips=False):
    """N2 fixtures: joint reconstruction (treeanc.py:934-1080) on the inputs of an existing fixture."""
    z = np.load(os.path.join(OUT, src + '.npz'))
    newick = str(z['newick'])
    aln = {str(k): np.array(list(str(v))) for k, v in zip(z['aln_names'], z['aln_seqs'])}
    tt = refenv.reference_treeanc(newick, aln, gtr, rng_seed=1)
    topo, flat, g = flatten_treeanc(tt)
    N1 = tt.infer_ancestral_sequences(marginal=False, reconstruct_tip_states=reconstruct_tips, debug=True)   # keep Lx / Cx
    out = dict(source=np.array(src), reconstruct_tips=np.array(reconstruct_tips), N_diff_first=np.array(N1),
               sequence_joint_LH=np.array(tt.tree.sequence_joint_LH), sequence_LH=np.array(tt.tree.sequence_LH),
               cseq=np.array([''.join(n.cseq) if n.cseq is not None else '' for n in topo.nodes]),
               root_joint_Lx=np.array(tt.tree.root.joint_Lx))
    stored = list(range(1, topo.n_nodes, max(1, topo.n_nodes // 8)))
    out['stored_nodes'] = np.array(stored)
    for i in stored:
        n = topo.nodes[i]
        if hasattr(n, 'joint_Lx') and getattr(n, 'joint_Cx', None) is not None:
            out['Lx_%d' % i] = np.array(n.joint_Lx)
            out['Cx_%d' % i] = np.array(n.joint_Cx)
    out['N_diff_second'] = np.array(tt.infer_ancestral_sequences(marginal=False, reconstruct_tip_states=reconstruct_tips))
    out['N_diff_marginal_after'] = np.array(tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=reconstruct_tips))
    if not reconstruct_tips:
        # joint branch-length optimisation (treeanc.py:1176-1243,1449-1473) from a fresh object
        t2 = refenv.reference_treeanc(newick, aln, gtr, rng_seed=1)
        t2.optimize_tree(branch_length_mode='joint', max_iter=2, prune_short=False)
        out['opt_joint_branch_length'] = np.array([n.branch_length for n in t2.tree.find_clades()])
        out['opt_joint_sequence_LH'] = np.array(t2.tree.unconstrained_sequence_LH)
    name = 'joint_' + src + ('_tips' if reconstruct_tips else '')
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-22s joint_LH=%.6f N_diff=%d/%d/%d %d KB' % (name, out['sequence_joint_LH'], N1, out['N_diff_second'],
                                                       out['N_diff_marginal_after'], os.path.getsize(path) // 1024))


def joint_cases():
    from treetime import GTR
    nuc = lambda: GTR.custom(pi=np.array([.3, .2, .2, .29, .01]), W=np.ones((5, 5)), alphabet='nuc')  # noqa: E731
    run_joint_case('nuc40', nuc())
    run_joint_case('nuc40', nuc(), reconstruct_tips=True)
    run_joint_case('poly70', nuc())
    run_joint_case('aa16_jtt92', GTR.standard('JTT92'), reconstruct_tips=True)


def run_extras_case(src, make_gtr):
    """Fixtures for the paths beyond the plain pass, on the inputs of an existing fixture: sampled sequences
    (sample_from_profile=True, treeanc.py:786-798,919-923), per-branch masks (ARG mode, :862-872,909-917,1294,1564-1572)
    and site-specific GTR inference (:1594-1627, gtr_site_specific.py:207-310)."""
    if ONLY and ('extras_' + src) not in ONLY:
        return
    z = np.load(os.path.join(OUT, src + '.npz'))
    newick = str(z['newick'])
    aln = {str(k): np.array(list(str(v))) for k, v in zip(z['aln_names'], z['aln_seqs'])}
    out = dict(source=np.array(src))
    # 1. sampled sequences, generator seeded with 7
    tt = refenv.reference_treeanc(newick, aln, make_gtr(), rng_seed=7)
    nodes = list(tt.tree.find_clades())
    for k, tips in enumerate((False, True)):
        out['sample_N_diff_%d' % k] = np.array(tt.infer_ancestral_sequences(marginal=True, sample_from_profile=True, reconstruct_tip_states=tips))
        out['sample_cseq_%d' % k] = np.array([''.join(n.cseq) if n.cseq is not None else '' for n in nodes])
    out['sample_next_uniform'] = np.array(tt.rng.random())
    # 2. per-branch masks: node k carries the segment mask if k % 3 == 0 else the all-ones mask
    tt = refenv.reference_treeanc(newick, aln, make_gtr(), rng_seed=1)
    nodes = list(tt.tree.find_clades())
    L = tt.data.compressed_length
    seg = np.zeros(L); seg[:L // 2] = 1
    for k, n in enumerate(nodes):
        n.mask = seg if k % 3 == 0 else np.ones(L)
    out['mask_segment'] = seg
    out['mask_N_diff'] = np.array(tt.infer_ancestral_sequences(marginal=True))
    out['mask_total_LH'] = np.array(tt.tree.total_sequence_LH)
    out['mask_sequence_LH'] = np.array(tt.tree.sequence_LH)
    out['mask_cseq'] = np.array([''.join(n.cseq) if n.cseq is not None else '' for n in nodes])
    bl_nodes = list(range(1, len(nodes), max(1, len(nodes) // 8)))[:8]
    out['mask_bl_nodes'] = np.array(bl_nodes)
    out['mask_bl_opt'] = np.array([tt.optimal_marginal_branch_length(nodes[i]) for i in bl_nodes])
    out['mask_profile_nodes'] = np.array([i for i in bl_nodes if not nodes[i].is_terminal()])
    for i in out['mask_profile_nodes']:
        out['mask_profile_%d' % i] = np.array(nodes[i].marginal_profile)
        out['mask_outgroup_%d' % i] = np.array(nodes[i].marginal_outgroup_LH)
    g = tt.infer_gtr(marginal=True, pc=1.0)
    out['mask_inferred_W'] = np.array(g.W); out['mask_inferred_Pi'] = np.array(g.Pi)
    # 3. site-specific GTR inference (no pattern compression), then a reconstruction under the inferred model
    tt = refenv.reference_treeanc(newick, aln, make_gtr(), rng_seed=1, compress=False)
    tt.infer_ancestral_sequences(marginal=True)
    g = tt.infer_gtr(marginal=True, site_specific=True, pc=1.0)
    out['ss_Pi'] = np.array(g.Pi); out['ss_mu'] = np.array(g.mu); out['ss_W'] = np.array(g.W)
    out['ss_N_diff'] = np.array(tt.infer_ancestral_sequences(marginal=True))
    out['ss_total_LH'] = np.array(tt.tree.total_sequence_LH)
    path = os.path.join(OUT, 'extras_' + src + '.npz')
    np.savez_compressed(path, **out)
    print('%-22s sampled N_diff=%d/%d  masked LH=%.6f  site-specific LH=%.6f  %d KB' % (
        'extras_' + src, out['sample_N_diff_0'], out['sample_N_diff_1'], out['mask_total_LH'], out['ss_total_LH'], os.path.getsize(path) // 1024))


def extras_cases():
    from treetime import GTR
    run_extras_case('nuc40', lambda: GTR.custom(pi=np.array([.3, .2, .2, .29, .01]), W=np.ones((5, 5)), alphabet='nuc'))


def chars(idx, gtr):
    return {k: gtr.alphabet[v] for k, v in idx.items()}


def main():
    from treetime import GTR
    from treetime.gtr_site_specific import GTR_site_specific
    from treetime.seqgen import SeqGen
    from Bio import Phylo
    from io import StringIO

    # 0. the reference's own known-answer test (test/test_treetime.py:140-151)
    nwk = '((A:0.60100000009,B:0.3010000009):0.1,C:0.2):0.001;'
    aln = {'A': np.array(list('AAAAAAAAAAAAAAAACCCCCCCCCCCCCCCCGGGGGGGGGGGGGGGGTTTTTTTTTTTTTTTT')),
           'B': np.array(list('AAAACCCCGGGGTTTTAAAACCCCGGGGTTTTAAAACCCCGGGGTTTTAAAACCCCGGGGTTTT')),
           'C': np.array(list('ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT'))}
    kat = GTR.custom(alphabet=np.array(['A', 'C', 'G', 'T']), pi=np.array([0.9, 0.06, 0.02, 0.02]), W=np.ones((4, 4)))
    run_case('kat3', nwk, aln, kat, n_bl=4)

    nuc = lambda: GTR.custom(pi=np.array([.3, .2, .2, .29, .01]), W=np.ones((5, 5)), alphabet='nuc')  # noqa: E731
    # 1. nucleotides with ambiguity codes; reference SeqGen makes the alignment
    tree = synth.random_tree(40, seed=11, mean_bl=0.01)
    sg = SeqGen(400, tree=Phylo.read(StringIO(tree.to_newick()), 'newick'), gtr=nuc(), rng_seed=11, verbose=0)
    sg.evolve()
    aln = synth.sprinkle_ambiguous({r.id: np.array(list(str(r.seq))) for r in sg.get_aln()}, 0.02, 'N-RY', seed=12)
    run_case('nuc40', tree.to_newick(), aln, nuc(), store_every=3, optimize=True, infer_gtr=True)
    run_case('nuc40_tips', tree.to_newick(), aln, nuc(), store_every=5, reconstruct_tips=True)
    # 2. polytomies, zero-length branches
    tree = synth.random_tree(70, seed=13, mean_bl=0.01, polytomy_frac=0.5, zero_frac=0.25)
    g = nuc()
    aln = chars(synth.evolve_alignment(tree, 300, g.Pi, g.W, seed=13), g)
    run_case('poly70', tree.to_newick(), aln, g, store_every=4, optimize=True)
    # 3. amino acids, the reference's JTT92 (20 states)
    g20 = GTR.standard('JTT92')
    tree = synth.random_tree(16, seed=14, mean_bl=0.05)
    aln = chars(synth.evolve_alignment(tree, 90, g20.Pi, g20.W, seed=14), g20)
    run_case('aa16_jtt92', tree.to_newick(), aln, g20, store_every=2)
    # 3b. 22-state 'aa' alphabet
    g22 = GTR.random(alphabet='aa', rng=np.random.default_rng(3))
    aln = chars(synth.evolve_alignment(tree, 90, g22.Pi, g22.W, seed=15), g22)
    run_case('aa16_q22', tree.to_newick(), aln, g22, store_every=3)
    # 4. site-specific model (interpolated expQt, the reference's default)
    gs = GTR_site_specific.random(L=120, alphabet='nuc', rng=np.random.default_rng(16))
    tree = synth.random_tree(20, seed=16, mean_bl=0.05)
    aln = chars(synth.evolve_alignment(tree, 120, gs.Pi.mean(axis=1), gs.W, seed=16), gs)
    run_case('sitespec20', tree.to_newick(), aln, gs, store_every=3, compress=False)
    # 5. BASELINE.json configs[0] at full size: inputs regenerated from seeds at test time
    tree = synth.random_tree(200, seed=1, mean_bl=2e-3)
    g = nuc()
    idx = synth.evolve_alignment(tree, 1400, g.Pi, g.W, seed=1)
    run_case('cfg1_200x1400', tree.to_newick(), chars(idx, g), g, store_every=50, store_inputs=False, optimize=True,
             extra=dict(input_sha=np.array(sha(np.vstack([idx[k] for k in sorted(idx)]))),
                        gen=np.array('synth.random_tree(200, seed=1, mean_bl=2e-3); synth.evolve_alignment(tree, 1400, Pi, W, seed=1)')))
    # 6. BASELINE.json configs[1] at full size
    if '--skip-big' not in sys.argv:
        tree = synth.random_tree(2000, seed=1, mean_bl=5e-4)
        idx = synth.evolve_alignment(tree, 10000, g.Pi, g.W, seed=1)
        run_case('cfg2_2000x10000', tree.to_newick(), chars(idx, g), g, store_every=100000, store_inputs=False, n_bl=4, cseq_as_idx=True,
                 extra=dict(input_sha=np.array(sha(np.vstack([idx[k] for k in sorted(idx)]))),
                            gen=np.array('synth.random_tree(2000, seed=1, mean_bl=5e-4); synth.evolve_alignment(tree, 10000, Pi, W, seed=1)')))


if __name__ == '__main__':
    if '--only-joint' in sys.argv:      # adds the joint_*.npz fixtures next to the existing ones
        joint_cases()
    elif '--only-extras' in sys.argv:   # adds the extras_*.npz fixtures (sampling, masks, site-specific inference)
        extras_cases()
    else:
        main()
        joint_cases()
        extras_cases()
