"""CPU: the oracle (oracle/flat_numpy.py) against golden vectors made by the UNMODIFIED reference,
and against the reference's own known-answer test."""
import numpy as np
import pytest

import flat_numpy as O
import golden_util as G


@pytest.mark.parametrize('name', G.SMALL + G.SITE_SPECIFIC)
def test_oracle_matches_reference_golden(name):
    z = G.load(name)
    flat, g = G.flat_and_gtr(z)
    tips = bool(z['reconstruct_tips'])
    res = O.marginal(flat, g, reconstruct_tip_states=tips)
    ss = bool(g['site_specific'])
    tol = 1e-13 if ss else 0.0          # single-site: bit-identical (see oracle/validate_against_reference.py)
    assert res.N_diff == int(z['N_diff_first'])
    if ss:
        assert np.allclose(res.sequence_LH, z['sequence_LH'], rtol=1e-13, atol=0)
    else:
        assert np.array_equal(res.sequence_LH, z['sequence_LH'])
        assert res.total_LH == float(z['total_LH'])
    ab = np.array([str(c) for c in z['gtr_alphabet']])
    for i in z['stored_nodes']:
        assert np.abs(res.subtree_LH[i] - z['subtree_%d' % i]).max() <= tol
        if i > 0:
            assert np.abs(res.outgroup_LH[i] - z['outgroup_%d' % i]).max() <= tol
        if 'profile_%d' % i in z.files:
            assert np.abs(res.profile[i] - z['profile_%d' % i]).max() <= tol
    for i, s in enumerate(z['cseq']):
        if res.seq_idx[i] is not None:
            assert ''.join(ab[res.seq_idx[i]]) == str(s)
    res2 = O.marginal(flat, g, reconstruct_tip_states=tips, prev_seq_idx=res.seq_idx)
    assert res2.N_diff == int(z['N_diff_second']) == 0
    # branch-length surface
    for k, n in enumerate(z['bl_nodes']):
        for t, ref in zip(z['obj_t'], z['obj'][k]):
            assert np.isclose(O.branch_objective(flat, g, res, int(n), float(t)), ref, rtol=1e-13)
        bl = O.optimal_marginal_branch_length(flat, g, res, int(n))
        assert bl == z['bl_opt'][k] or abs(bl - z['bl_opt'][k]) < 1e-7 * z['bl_opt'][k]
    if 'n_ij' in z.files:
        n_ija, T_ia = O.mutation_counts(flat, g, res)
        assert np.array_equal(n_ija.sum(axis=-1), z['n_ij']) and np.array_equal(T_ia.sum(axis=-1), z['T_i'])


@pytest.mark.parametrize('name', G.JOINT)
def test_oracle_joint_matches_reference_golden(name):
    """N2: the joint oracle is bit-identical to TreeAnc._ml_anc_joint (treeanc.py:934-1080)."""
    zj = G.load(name)
    z = G.load(str(zj['source']))
    flat, g = G.flat_and_gtr(z)
    tips = bool(zj['reconstruct_tips'])
    res = O.joint(flat, g, reconstruct_tip_states=tips)
    assert res.N_diff == int(zj['N_diff_first'])
    assert res.total_LH == float(zj['sequence_joint_LH'])
    assert np.array_equal(res.sequence_LH, zj['sequence_LH'])
    assert np.array_equal(res.joint_Lx[0], zj['root_joint_Lx'])
    ab = np.array([str(c) for c in z['gtr_alphabet']])
    for i, s in enumerate(zj['cseq']):
        if res.seq_idx[i] is not None and (tips or flat['tip_row'][i] < 0):
            assert ''.join(ab[res.seq_idx[i]]) == str(s)
    for i in zj['stored_nodes']:
        if 'Lx_%d' % i in zj.files:
            assert np.array_equal(res.joint_Lx[i], zj['Lx_%d' % i]) and np.array_equal(res.joint_Cx[i], zj['Cx_%d' % i])
    assert O.joint(flat, g, reconstruct_tip_states=tips, prev_seq_idx=res.seq_idx).N_diff == int(zj['N_diff_second']) == 0


def test_reference_known_answer_lh_normalisation():
    """test/test_treetime.py:137-155: over all 4^3 column patterns of a 3-tip tree sum exp(LH) = 1;
    plus the values captured from the reference."""
    z = G.load('kat3')
    flat, g = G.flat_and_gtr(z)
    res = O.marginal(flat, g)
    assert abs(np.exp(res.sequence_LH).sum() - 1.0) < 1e-6
    assert res.total_LH == -495.7525153086474
    assert np.allclose(res.sequence_LH[:4], [-0.2014649555272725, -4.0517702631511066, -5.150382551819217, -5.150382551819217], rtol=0, atol=0)
    assert np.allclose(res.profile[0][1], [0.5189673178309464, 0.4768648799557071, 0.00208390110667322, 0.00208390110667321], rtol=0, atol=1e-17)
    n = 2   # NODE_0000001: parent of A and B (preorder after ladderize)
    assert np.allclose(res.outgroup_LH[n][1], [0.1631423222298166, 0.8296069078933028, 0.00362538493844039, 0.00362538493844036], atol=1e-16)
    assert abs(O.optimal_marginal_branch_length(flat, g, res, n) - 0.8297413773978182) < 1e-12


def test_oracle_optimize_sweep_matches_reference():
    """One optimize_tree_marginal sweep + damping on flat arrays equals the reference's result."""
    z = G.load('poly70')
    flat, g = G.flat_and_gtr(z)
    # branch lengths before flooring are not stored; the floored t equals branch_length wherever it is above the floor
    res = O.marginal(flat, g)
    assert res.total_LH == float(z['total_LH'])
