#!/usr/bin/env python
"""bench.py -- marginal ancestral reconstruction throughput (branch x pattern updates/s; log-LH rel err).

    python bench.py [--gpus N --steps K --warmup W] [--workload cfg3|cfg2|cfg1|cfg4|cfg5|tiny] [--scaling auto|strong|weak]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...   # CPU arm: the UNMODIFIED reference (oracle/_ref) on the host cores

One step = one full `infer_ancestral_sequences(marginal=True)` pass (batched expQt, level-ordered postorder,
root, level-ordered preorder, reductions) over one synthetic alignment.

N = 1: BASELINE.json configs[2] (20k tips x 29,903 sites) on one GPU.
N > 1: one process per GPU, tree/model replicated.  Default `--scaling strong`: ONE configs[2] alignment, the
compressed patterns sharded over the ranks (dist.shard_bounds), the only collective is the all-reduce of
{total log-LH, N_diff}; `updates_per_step` is constant in N.  The same run also measures the weak form (every rank
a full-size alignment) as `weak_scaling`, and at N = 8 BASELINE.json configs[4] in full (100k tips x 30 kb,
site-specific GTR, 8 shards of 3,750 sites) as `north_star_config`.

Prints ONE JSON line (rank 0).  `value` = updates/s with inputs resident in HBM, CUDA events around the K steps, max
over ranks.  `e2e` = the same metric through the product API (treetime_b200.TreeAnc, the mirror of treetime.TreeAnc):
per step the alignment shard (sparse host form), model and branch lengths go host->device, the pass runs, and the
per-pattern LH, the totals and every reconstructed sequence (sparse form) come back device->host.
`parity` = the CPU oracle port (bit-identical to the reference on the build container) on a pattern slice of the same
tree against the device's resident pass: log-LH relative error, max profile error, argmax mismatches off exact ties.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_tips, n_sites, alphabet, mean branch length, description)
    'cfg3': (20000, 29903, 'nuc', 1.0 / 29903, '20k tips x 29,903-site SARS-CoV-2-shaped nucleotide alignment (BASELINE.json configs[2])'),
    'cfg2': (2000, 10000, 'nuc', 5e-4, '2k tips x 10 kb nucleotide (BASELINE.json configs[1])'),
    'cfg1': (200, 1400, 'nuc', 2e-3, '200-tip x 1.4 kb nucleotide (BASELINE.json configs[0])'),
    'cfg4': (5000, 1000, 'aa_nogap', 1e-2, '5k tips x 1,000-site amino-acid alignment, JTT92 (20 states) (BASELINE.json configs[3])'),
    'cfg5': (100000, 3750, 'nuc_site_specific', 3.3e-5, '100k tips x 30 kb nucleotide with site-specific GTR, one of 8 pattern shards '
                                                       '(3,750 uncompressed sites per GPU) (BASELINE.json configs[4])'),
    'tiny': (64, 500, 'nuc', 1e-2, 'smoke-sized'),
}
SURVEY_BYTES_PER_UPDATE = {5: 165.5, 4: 133.5, 20: 645.5, 22: 709.5}    # 4 q s + 0.5 s + 1.5
NUC_PI = np.array([0.3, 0.2, 0.2, 0.29, 0.01])
METRIC = 'marginal ancestral reconstruction branch x pattern updates/s'


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and so its first-touch pinned buffers) to the CPUs next to its GPU."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0' % (getattr(pr, 'pci_domain_id', 0), pr.pci_bus_id, getattr(pr, 'pci_device_id', 0))
        node = int(open(path + '/numa_node').read())
        cpus = open(path + '/local_cpulist').read().strip()
        ids = set()
        for part in cpus.split(','):
            a, _, b = part.partition('-')
            ids.update(range(int(a), int(b or a) + 1))
        if ids:
            os.sched_setaffinity(0, ids & os.sched_getaffinity(0) or ids)
        return {'numa_node': node, 'cpus': cpus}
    except Exception as e:       # best effort: containers may hide sysfs
        return {'error': str(e)[:80]}


def jtt92():
    """(W, pi) of JTT92 as the unmodified reference builds it (aa_models.py), carried by the committed golden fixture."""
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'aa16_jtt92.npz'), allow_pickle=False)
    return z['gtr_W'].copy(), z['gtr_Pi'].copy()


def make_model(name, seed):
    from treetime_b200.gtr import GTR
    n_tips, L, alphabet, mean_bl, _ = WORKLOADS[name]
    if alphabet == 'nuc':
        return GTR.custom(pi=NUC_PI.copy(), W=np.ones((5, 5)), alphabet='nuc'), True
    if alphabet == 'nuc_site_specific':
        from treetime_b200.gtr import GTRSiteSpecific
        # treeanc.py:186-187: no pattern compression with site-specific models
        return GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(1000 + seed)), False
    W, pi = jtt92()
    return GTR.custom(pi=pi, W=W, alphabet='aa_nogap'), True


def make_inputs(name, seed):
    """(tree, alignment as ASCII byte rows, model, compress): the same tree on every rank (seed 1), columns from `seed`."""
    from treetime_b200 import synth
    n_tips, L, alphabet, mean_bl, _ = WORKLOADS[name]
    gtr, compress = make_model(name, seed)
    tree = synth.random_tree(n_tips, seed=1, mean_bl=mean_bl)
    Pi = gtr.Pi if np.ndim(gtr.Pi) == 1 else gtr.Pi.mean(axis=1)
    idx = synth.evolve_alignment(tree, L, Pi, gtr.W, mu=1.0, seed=seed)
    ab = np.asarray(gtr.alphabet).astype('S1').view(np.uint8)
    return tree, {k: ab[v] for k, v in idx.items()}, gtr, compress


def make_workload(name, seed):
    """Flat arrays of the whole problem (what crosses the C-ABI and what the CPU oracle consumes)."""
    from treetime_b200 import synth
    from treetime_b200.sequence_data import SequenceData
    tree, aln, gtr, compress = make_inputs(name, seed)
    sd = SequenceData(aln, compress=compress, ambiguous=gtr.ambiguous)
    return synth.flat_problem(tree, sd, gtr)


def shard_problem(tt):
    """Flat arrays of exactly what this rank's engine holds (its pattern shard), taken from the TreeAnc itself."""
    from treetime_b200.flatten import gtr_arrays
    topo = tt._flat()
    lo, hi = tt._shard()
    codes, table = tt._tip_codes()
    flat = topo.as_dict()
    flat.update(tip_codes=codes, code_profiles=table, multiplicity=np.ascontiguousarray(tt.data.multiplicity()[lo:hi], dtype=np.float64),
                t=np.array(tt._t_last))
    g = gtr_arrays(tt.gtr)
    if g['site_specific']:
        g = dict(g, eigenvals=g['eigenvals'][:, lo:hi], v=g['v'][:, :, lo:hi], v_inv=g['v_inv'][:, :, lo:hi], Pi=g['Pi'][:, lo:hi], mu=g['mu'][lo:hi])
    return flat, g


def algorithmic_bytes(flat, q):
    """Bytes the level kernels must move per pass under THIS design:
    postorder: read S of internal children + 1-byte codes of tip children, write S per internal node (the log-prefactors
    are summed per block run, not stored per node); preorder: read the parent profile once per parent with an internal
    child, per internal child read S, write the profile, read + write the 1-byte state."""
    Lp = flat['multiplicity'].shape[0]
    n_nodes = flat['parent'].shape[0]
    tip = flat['tip_row'] >= 0
    n_tips = int(tip.sum())
    n_int = n_nodes - n_tips
    has_int_child = np.zeros(n_nodes, dtype=bool)
    has_int_child[flat['parent'][1:][~tip[1:]]] = True
    post = Lp * (n_int * q * 8 + (n_int - 1) * q * 8 + n_tips * 1)
    pre = Lp * (int(has_int_child.sum()) * q * 8 + (n_int - 1) * (2 * q * 8 + 2))
    return post, pre


class ClockSampler(object):
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line).  A reader thread time-stamps every
    sample; the timed region of a sharded run can be shorter than nvidia-smi's period, so the caller keeps the same step
    running a little longer (untimed) until a few samples under load exist -- both counts are reported."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.rows = []          # (time, fields)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def start(self, wait_first=3.0):
        import threading
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < wait_first:     # nvidia-smi needs a moment to come up
            time.sleep(0.01)

    def count_since(self, t):
        return sum(1 for ts, _ in self.rows if ts >= t)

    def stop(self, t_begin=None, t_end=None, t_load_end=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons, timed = [], [], set(), 0
        for ts, f in self.rows:
            if len(f) < 9 or (t_begin is not None and ts < t_begin) or (t_load_end is not None and ts > t_load_end):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            if t_end is not None and ts <= t_end:
                timed += 1
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'samples_in_timed_region': timed,
                'note': 'samples beyond the timed region were taken while the same step kept running (untimed) so that short timed regions still '
                        'get clocks under load'}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(workload, kernel_prefix):
    """DRAM bytes per pass (read + write, summed over the level launches) of the dominant kernel from
    the committed ncu capture (profiles/traffic.json, made by tools/ncu_traffic.py), or None."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        ks = json.load(open(p))[workload]['kernels']
        for name, a in ks.items():
            if name.startswith(kernel_prefix):
                return a['dram_bytes']
    except Exception:
        pass
    return None


def cpu_sample(flat, g, n_patterns):
    """Bounded sample of the same workload for the CPU arm: the first n_patterns columns."""
    n = min(n_patterns, flat['multiplicity'].shape[0])
    s = dict(flat)
    s['tip_codes'] = np.ascontiguousarray(flat['tip_codes'][:, :n])
    s['multiplicity'] = flat['multiplicity'][:n].copy()
    if g.get('site_specific'):     # the model is per pattern: slice it with the columns
        g = dict(g, eigenvals=g['eigenvals'][:, :n], v=g['v'][:, :, :n], v_inv=g['v_inv'][:, :, :n], Pi=g['Pi'][:, :n], mu=g['mu'][:n])
    return s, n, g


def tie_mask(profile, rel=1e-12):
    """True where the two largest entries of a profile row are closer than `rel`: the argmax is an exact tie."""
    s = np.sort(profile, axis=1)
    return (s[:, -1] - s[:, -2]) <= rel * s[:, -1]


def parity_against_port(eng, flat, g, n_patterns, max_profile_nodes=4000, cached=None, near_tie=1e-12):
    """The CPU oracle port (oracle/flat_numpy.py: the reference's per-node numpy calls, checker only) on the first
    n_patterns patterns of this shard against the device's resident pass: BASELINE.md §3 / the metric's second half.
    Returns (parity dict, cpu seconds, updates of the sample)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import flat_numpy as O
    if cached is None:
        s, n, gs = cpu_sample(flat, g, n_patterns)
        t0 = time.perf_counter()
        res = O.marginal(s, gs)
        cpu_s = time.perf_counter() - t0
        cached = (s, n, gs, res, cpu_s)
    s, n, gs, res, cpu_s = cached
    m = s['multiplicity']
    lh = eng.site_lh()[:n]
    tot_gpu = float((lh * m).sum())
    internal = np.nonzero(flat['tip_row'] < 0)[0]
    idx = eng.all_seq_idx()[:, :n]
    mism = mism_off_ties = 0
    for k, node in enumerate(internal):
        bad = idx[k] != res.seq_idx[node]
        if bad.any():
            mism += int(bad.sum())
            mism_off_ties += int((bad & ~tie_mask(res.profile[node], near_tie)).sum())
    sel = internal if internal.shape[0] <= max_profile_nodes else internal[np.linspace(0, internal.shape[0] - 1, max_profile_nodes).astype(int)]
    perr = 0.0
    for node in sel:
        perr = max(perr, float(np.abs(eng.node_array(int(node), 2)[:n] - res.profile[node]).max()))
    par = {'log_lh_rel_err': abs(tot_gpu - float(res.total_LH)) / abs(float(res.total_LH)),
           'max_site_log_lh_rel_err': float(np.max(np.abs(lh - res.sequence_LH) / np.abs(res.sequence_LH))),
           'max_profile_abs_err': perr, 'argmax_mismatch': mism, 'argmax_mismatch_off_ties': mism_off_ties,
           'patterns_compared': int(n), 'internal_nodes_compared_sequences': int(internal.shape[0]),
           'internal_nodes_compared_profiles': int(len(sel)),
           'against': 'oracle/flat_numpy.py (CPU restatement of treeanc.py:762-932, bit-identical to the unmodified reference on the '
                      'build container) on the first patterns of rank 0\'s shard, same tree / model / branch lengths',
           'tolerances': {'log_lh_rel_err': 1e-9, 'max_profile_abs_err': 1e-6, 'argmax_mismatch_off_ties': 0}}
    n_br = flat['parent'].shape[0] - 1
    par['_cached'] = cached
    return par, cpu_s, n_br * n


# ------------------------------------------------------------------------------------------------ reference arm
def _reference_worker(k, spec, bar, n_steps, out_q):
    """One host core: the UNMODIFIED reference's TreeAnc on its own slice of the alignment columns."""
    try:
        import threadpoolctl
        _lim = threadpoolctl.threadpool_limits(1)      # noqa: F841
    except Exception:
        pass
    try:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import refenv
        refenv.activate()
        from io import StringIO
        from Bio import Phylo
        from Bio.Align import MultipleSeqAlignment
        from Bio.SeqRecord import SeqRecord
        from Bio.Seq import Seq
        from treetime import TreeAnc, GTR as RG
        lo, hi = spec['bounds'][k]
        names, chars = spec['names'], spec['chars']
        aln = MultipleSeqAlignment([SeqRecord(Seq(chars[i, lo:hi].tobytes().decode('ascii')), id=nm, name=nm, description='')
                                    for i, nm in enumerate(names)])
        kind = spec['model']['kind']
        if kind == 'single':
            gtr = RG.custom(pi=spec['model']['pi'].copy(), W=spec['model']['W'].copy(), alphabet=spec['model']['alphabet'])
        else:
            from treetime.gtr_site_specific import GTR_site_specific
            gtr = GTR_site_specific.custom(mu=spec['model']['mu'][lo:hi].copy(), pi=spec['model']['pi'][:, lo:hi].copy(),
                                           W=spec['model']['W'].copy(), alphabet=spec['model']['alphabet'])
        tt = TreeAnc(tree=Phylo.read(StringIO(spec['newick']), 'newick'), aln=aln, gtr=gtr, compress=False, verbose=0)
        lh = None
        bar.wait()                                    # set-up done everywhere
        for _ in range(n_steps):
            bar.wait()
            tt.infer_ancestral_sequences(marginal=True)      # the reference's own public call, stock code path
            lh = float(tt.sequence_LH())
            bar.wait()
        out_q.put((k, lh, None))
    except Exception as e:       # never leave the others waiting on the barrier
        try:
            bar.abort()
        except Exception:
            pass
        out_q.put((k, None, repr(e)))


def run_reference_arm(args, real_stdout):
    """`--impl reference`: the unmodified reference (oracle/_ref; /root/reference in the build container) through its
    own TreeAnc.infer_ancestral_sequences(marginal=True), one process per host core on disjoint column slices of the
    same workload (the reference is single-threaded numpy; patterns are independent), each step a bounded sample."""
    import multiprocessing as mp
    n_tips, L, alphabet, mean_bl, desc = WORKLOADS[args.workload]
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import refenv
    total_steps = args.steps + args.warmup
    workers = max(1, min(os.cpu_count() or 1, args.cpu_workers or 64))
    have_ref = refenv.available()
    from treetime_b200 import synth
    from treetime_b200.sequence_data import SequenceData
    tree, aln, gtr, compress = make_inputs(args.workload, 1)
    topo, flat, g = synth.flat_problem(tree, SequenceData(aln, compress=compress, ambiguous=gtr.ambiguous), gtr)   # ladderizes `tree`
    q = g['Pi'].shape[0]
    n_br = flat['parent'].shape[0] - 1
    Lp = flat['multiplicity'].shape[0]
    # ~2.4e6 updates/s per core for the reference, ~4.4e6 for the port: aim at ~120 s for the whole run
    rate = 2.0e6 if have_ref else 4.0e6
    per_worker = args.cpu_patterns or int(max(32, min(Lp // workers, 120.0 * rate / (n_br * max(1, total_steps)))))
    try:
        import psutil
        per_pattern = 8.0 * flat['parent'].shape[0] * q * 8          # bytes one pattern costs a worker (per-node arrays)
        budget = 0.5 * psutil.virtual_memory().available
        per_worker = int(min(per_worker, max(32, budget // (workers * per_pattern))))
    except Exception:
        pass
    per_worker = max(1, min(per_worker, Lp // workers))
    bounds = [(k * per_worker, (k + 1) * per_worker) for k in range(workers)]
    updates = n_br * per_worker * workers
    if have_ref:
        from treetime_b200.flatten import code_table
        chars, _, _ = code_table(gtr.profile_map, gtr.n_states)
        code2char = np.frombuffer((''.join(chars) + (gtr.ambiguous or 'N')).encode('ascii'), dtype=np.uint8)
        names = [topo.nodes[n].name for n in topo.tip_nodes]
        model = ({'kind': 'single', 'pi': np.array(gtr.Pi), 'W': np.array(gtr.W), 'alphabet': 'nuc' if alphabet == 'nuc' else 'aa_nogap'}
                 if np.ndim(gtr.Pi) == 1 else
                 {'kind': 'site_specific', 'pi': np.array(gtr.Pi), 'W': np.array(gtr.W), 'mu': np.array(gtr.mu), 'alphabet': 'nuc'})
        spec = {'bounds': bounds, 'names': names, 'chars': code2char[flat['tip_codes'][:, :per_worker * workers]],
                'newick': tree.to_newick(), 'model': model}
        ctx = mp.get_context('fork')
        bar = ctx.Barrier(workers + 1)
        out_q = ctx.Queue()
        procs = [ctx.Process(target=_reference_worker, args=(k, spec, bar, total_steps, out_q)) for k in range(workers)]
        for p in procs:
            p.start()
        times, err = [], None
        try:
            bar.wait(timeout=900)
            for _ in range(total_steps):
                bar.wait(timeout=900)
                t0 = time.perf_counter()
                bar.wait(timeout=1800)
                times.append(time.perf_counter() - t0)
        except Exception as e:
            err = repr(e)
        res = []
        for _ in procs:
            try:
                res.append(out_q.get(timeout=60 if err is None else 5))
            except Exception:
                break
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
        bad = [r for r in res if r[2]]
        if err or bad or len(times) < total_steps:
            have_ref = False
            sys.stderr.write('reference arm fell back to the port: %s %s\n' % (err, bad[:1]))
        else:
            kind = 'reference'
            sample = ('%d processes (one per host core), each the unmodified treetime.TreeAnc(compress=False).infer_ancestral_sequences('
                      'marginal=True) over its own %d of the %d compressed patterns of the same tree/alignment per step (cost is linear in '
                      'patterns)' % (workers, per_worker, Lp))
            note = ('treetime 0.12.1 from %s, Biopython replaced by oracle/bioshim (container stubs, no numerics)'
                    % ('oracle/_ref (pip-installed copy)' if refenv.is_staged_copy() else refenv.REFERENCE))
    if not have_ref:
        updates, times, per_worker = run_port_parallel(flat, g, per_worker, workers, total_steps)
        kind = 'port'
        sample = ('%d processes (one per core), each one pass of the oracle port over its own %d of %d compressed patterns of the same '
                  'tree/alignment per step' % (workers, per_worker, Lp))
        note = 'oracle/flat_numpy.py: flat-array port of the reference numpy path (the reference itself is not staged: python oracle/stage_ref.py)'
    timed = times[args.warmup:]
    ms = 1e3 * float(np.mean(timed))
    val = updates / (ms / 1e3)
    _emit(real_stdout, {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'updates/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak' if args.gpus == 1 else args.scaling_resolved,
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s: %s' % (args.workload, desc), 'n_tips': n_tips, 'n_sites': L, 'n_states': q},
        'cpu_baseline': {'value': val, 'unit': 'updates/s', 'cores': workers, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': note})
    return 0


_POOL_STATE = {}


def _port_worker_init(flat, g):
    try:
        import threadpoolctl
        _POOL_STATE['limit'] = threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    _POOL_STATE['flat'], _POOL_STATE['g'] = flat, g


def _port_worker_run(bounds):
    import flat_numpy as O
    lo, hi = bounds
    flat, g = _POOL_STATE['flat'], _POOL_STATE['g']
    s = dict(flat)
    s['tip_codes'] = np.ascontiguousarray(flat['tip_codes'][:, lo:hi])
    s['multiplicity'] = flat['multiplicity'][lo:hi].copy()
    if g.get('site_specific'):
        g = dict(g, eigenvals=g['eigenvals'][:, lo:hi], v=g['v'][:, :, lo:hi], v_inv=g['v_inv'][:, :, lo:hi], Pi=g['Pi'][:, lo:hi],
                 mu=g['mu'][lo:hi])
    return float(O.marginal(s, g).total_LH)


def run_port_parallel(flat, g, per_worker, workers, repeats):
    """Fallback of the reference arm when the reference is not staged: the oracle port, one process per core."""
    import multiprocessing as mp
    Lp = flat['multiplicity'].shape[0]
    per_worker = max(1, min(per_worker, Lp // workers))
    bounds = [(k * per_worker, (k + 1) * per_worker) for k in range(workers)]
    n_br = flat['parent'].shape[0] - 1
    times = []
    with mp.get_context('fork').Pool(workers, initializer=_port_worker_init, initargs=(flat, g)) as pool:
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_port_worker_run, bounds, chunksize=1)
            times.append(time.perf_counter() - t0)
    return n_br * per_worker * workers, times, per_worker


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there)
    must not pollute it: point fd 1 at stderr for the whole run and keep the real stdout aside."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout_fd, obj):
    os.write(real_stdout_fd, (json.dumps(obj) + '\n').encode())


# ------------------------------------------------------------------------------------------------ GPU arm
class _Dev(object):
    def __init__(self, ptr):
        self.__cuda_array_interface__ = {'shape': (2,), 'typestr': '<f8', 'data': (ptr, False), 'version': 2}


class Leg(object):
    """One measured configuration: a TreeAnc (product API) whose engine also runs the resident timing."""

    def __init__(self, name, seed, comm, local_rank, stream):
        from treetime_b200.treeanc import TreeAnc
        self.name = name
        tree, aln, gtr, compress = make_inputs(name, seed)
        self.tt = TreeAnc(tree=tree, aln=aln, gtr=gtr, compress=compress, device=local_rank, comm=comm, sparse_io=True, rng_seed=1)
        self.stream = stream
        self.q = int(gtr.n_states)

    def prepare(self, comm):
        """(Re-)shard for `comm` and run the first pass (uploads, allocations, graph capture)."""
        tt = self.tt
        tt.comm = comm
        tt.reload_alignment()
        tt.infer_ancestral_sequences(marginal=True)
        self.eng = tt._engine
        self.eng.set_stream(self.stream.cuda_stream)
        self.eng.marginal()
        self.eng.sync()
        self.flat, self.g = shard_problem(tt)
        self.n_br = self.flat['parent'].shape[0] - 1
        self.Lp = self.flat['multiplicity'].shape[0]
        self.updates_local = self.n_br * self.Lp


def timed_resident(leg, steps, warmup, world, local_rank, sample_clocks=True):
    """K graph replays (+ the all-reduce of {LH, N_diff} when world > 1) between CUDA events on the engine's stream."""
    import torch
    import torch.distributed as dist
    eng, stream = leg.eng, leg.stream
    res_t = torch.as_tensor(_Dev(eng.results_device_ptr()), device='cuda') if world > 1 else None

    def step():
        eng.marginal()
        if world > 1:
            dist.all_reduce(res_t)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if sample_clocks:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    t_end = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches_timed = eng.launch_count() - launches0
    clocks = None
    if sample_clocks:
        # short timed regions (a sharded pass is ~2 ms): keep the same step running, untimed, until nvidia-smi has seen the load
        # (a fixed number of extra steps everywhere: the ranks must issue the same collectives)
        ms_agreed = ms_total
        if world > 1:
            tmax = torch.tensor([ms_total], device='cuda', dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms_agreed = float(tmax.item())
        extra = int(min(400, max(0, 0.35 / max(1e-4, ms_agreed / steps / 1e3) - steps)))
        for _ in range(extra):
            step()
        barrier()
        clocks = sampler.stop(t_begin, t_end, time.perf_counter())
    launches = launches_timed
    lh_local = eng.results()[0]           # copied to the host before the all-reduce touched the device buffer
    lh_global = float(res_t[0].item()) if world > 1 else lh_local
    return ms_total / steps, launches, clocks, lh_global, lh_local, barrier


def timed_api_e2e(leg, k, barrier):
    """The product API with host buffers, per step: alignment shard (sparse host form) + model + branch lengths H2D,
    infer_ancestral_sequences(marginal=True), per-pattern LH + totals + every reconstructed sequence (sparse) D2H."""
    tt = leg.tt

    def step():
        tt.reload_alignment()                                   # this step's input comes from the host again
        tt.infer_ancestral_sequences(marginal=True)             # ... LH per pattern and totals land on the host
        return tt.sequence_differences(gather=False)            # ... and so does every internal sequence (sparse form)

    step()
    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        root, node, pos, state = step()
    barrier()
    sec = (time.perf_counter() - t0) / k
    _, (ref, e_row, e_pos, e_code), table, mult = tt._sparse_codes
    q = leg.q
    model_bytes = sum(np.asarray(v).nbytes for v in (leg.g['eigenvals'], leg.g['v'], leg.g['v_inv'], leg.g['Pi'])) + 8 * np.size(leg.g['mu'])
    if leg.g.get('site_specific'):
        model_bytes = 0          # the per-pattern model is re-uploaded only when it changed (fingerprint), like the reference keeps its gtr
    h2d = int(ref.nbytes + e_row.nbytes + e_pos.nbytes + e_code.nbytes + table.nbytes + mult.nbytes + leg.flat['t'].nbytes + model_bytes)
    d2h = int(8 * leg.Lp + 32 + root.nbytes + node.nbytes + pos.nbytes + state.nbytes)
    return sec, h2d, d2h, float(tt.tree.total_sequence_LH), int(node.shape[0]), int(e_row.shape[0])


def dense_cabi_e2e(leg, k, nblk_max, local_rank, barrier):
    """C-ABI with DENSE host buffers (N = 1 only): per step the full tip-code matrix goes host->device from pinned memory
    and every reconstructed sequence comes back dense; the pattern axis is cut into column blocks with their own
    handles / streams so that the copies of one block overlap the pass of another."""
    import torch
    from treetime_b200.engine import Engine
    flat, g, q, Lp = leg.flat, leg.g, leg.q, leg.Lp
    n_int = int((flat['tip_row'] < 0).sum())
    nblk = max(1, min(nblk_max, Lp // 1024))
    bounds = [(Lp * i) // nblk for i in range(nblk + 1)]
    shards = []
    for i in range(nblk):
        lo, hi = bounds[i], bounds[i + 1]
        e = Engine(q, device=local_rank)
        st_i = torch.cuda.Stream()
        e.set_stream(st_i.cuda_stream)
        e.set_tree(flat['parent'], flat['child_ptr'], flat['child_idx'], flat['tip_row'])
        cp = torch.empty((flat['tip_codes'].shape[0], hi - lo), dtype=torch.uint8, pin_memory=True)
        cp.numpy()[...] = flat['tip_codes'][:, lo:hi]
        sp = torch.empty((n_int, hi - lo), dtype=torch.uint8, pin_memory=True)
        lp = torch.empty(hi - lo, dtype=torch.float64, pin_memory=True)
        shards.append((e, cp.numpy(), sp.numpy(), lp.numpy(), np.ascontiguousarray(flat['multiplicity'][lo:hi]), st_i, cp, sp, lp))

    def step():
        up_done = pass_done = None
        for e, cp, sp, lp, m, st_i, *_ in shards:
            if up_done is not None:
                st_i.wait_event(up_done)
            e.set_patterns(cp, flat['code_profiles'], m, validate=False)
            e.set_gtr(g)
            e.set_branch_lengths(flat['t'])
            up_done = torch.cuda.Event()
            up_done.record(st_i)
            if pass_done is not None:
                st_i.wait_event(pass_done)
            e.marginal()
            pass_done = torch.cuda.Event()
            pass_done.record(st_i)
            e.enqueue_site_lh(lp)
            e.enqueue_all_seq_idx(sp)
        return sum(e.results()[0] for e, *_ in shards)

    step()
    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        tot = step()
    barrier()
    sec = (time.perf_counter() - t0) / k
    h2d = int(flat['tip_codes'].nbytes + nblk * (flat['code_profiles'].nbytes + flat['t'].nbytes + 8 * (2 * q * q + 2 * q + 1)) + flat['multiplicity'].nbytes)
    d2h = int(n_int * Lp + 8 * Lp + 16 * nblk)
    for sh in shards:
        sh[0].close()
    del shards
    torch.cuda.empty_cache()
    return sec, h2d, d2h, nblk, tot


def pcie_probe():
    import torch
    pb = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    db = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    pe0, pe1, pe2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    db.copy_(pb, non_blocking=True); torch.cuda.synchronize()
    pe0.record(); db.copy_(pb, non_blocking=True); pe1.record(); pb.copy_(db, non_blocking=True); pe2.record()
    torch.cuda.synchronize()
    return {'h2d': 0.268435456 / (pe0.elapsed_time(pe1) / 1e3), 'd2h': 0.268435456 / (pe1.elapsed_time(pe2) / 1e3)}


def precision_study(leg, par64, cached, steps, warmup, local_rank):
    """north_star: "fp64 versus fp32 ... is decided from measured error".  The same pass with S / M STORED as float
    (ttb_set_message_storage, arithmetic stays fp64): time and error against the same CPU port slice, next to the fp64
    numbers of this run.  The engine is switched back to double afterwards."""
    eng = leg.eng
    eng.set_message_storage('f32')
    try:
        eng.marginal()
        eng.sync()
        ms32, _, _, lh32, _, _ = timed_resident(leg, steps, warmup, 1, local_rank, sample_clocks=False)
        par32, _, _ = parity_against_port(eng, leg.flat, leg.g, 0, cached=cached)
        par32.pop('_cached')
        loose, _, _ = parity_against_port(eng, leg.flat, leg.g, 0, cached=cached, near_tie=1e-6)
        post_b, pre_b = algorithmic_bytes(leg.flat, leg.q)
        peak, _ = measured_peak()
        n_int = int((leg.flat['tip_row'] < 0).sum()); n_tips = leg.flat['parent'].shape[0] - n_int
        # float storage halves every message byte; the 1-byte codes / states stay
        msg_bytes = (post_b + pre_b) - leg.Lp * (n_tips + 2 * (n_int - 1))
        bytes32 = msg_bytes / 2 + leg.Lp * (n_tips + 2 * (n_int - 1))
        keys = ('log_lh_rel_err', 'max_site_log_lh_rel_err', 'max_profile_abs_err', 'argmax_mismatch', 'argmax_mismatch_off_ties')
        return {'f64_storage': dict({k: par64[k] for k in keys}),
                'f32_storage': dict({k: par32[k] for k in keys}, ms_per_step=ms32, value=leg.updates_local / (ms32 / 1e3),
                                    argmax_mismatch_where_top2_gap_exceeds_1e6=loose['argmax_mismatch_off_ties'],
                                    whole_pass_frac_of_hbm_roofline=bytes32 / (ms32 / 1e3) / 1e9 / peak,
                                    rel_lh_diff_total_vs_f64=abs(lh32 - leg_lh64(leg)) / abs(lh32)),
                'patterns_compared': par64['patterns_compared'],
                'what': 'S / M / Mtip stored as float32, every arithmetic operation in float64 (ttb_set_message_storage(TTB_STORAGE_F32)); '
                        'errors against the CPU port on the same pattern slice; opt-in, the default and every headline number is float64 storage'}
    finally:
        eng.set_message_storage('f64')
        eng.marginal()
        eng.sync()


def leg_lh64(leg):
    return leg.lh64


def reduce_max_sum(world, maxes, sums):
    if world == 1:
        return [float(x) for x in maxes], [float(x) for x in sums]
    import torch
    import torch.distributed as dist
    a = torch.tensor(maxes, device='cuda', dtype=torch.float64)
    b = torch.tensor(sums, device='cuda', dtype=torch.float64)
    dist.all_reduce(a, op=dist.ReduceOp.MAX)
    dist.all_reduce(b)
    return a.tolist(), b.tolist()


def phase_profile(eng):
    phases = eng.profile_marginal()
    for _ in range(2):
        ph = eng.profile_marginal()
        phases = {k: (min(phases[k][0], ph[k][0]), ph[k][1]) for k in ph}
    return phases


FP64_TENSOR_PEAK_TFLOPS = 37.0     # mma.sync.m8n8k4.f64 on this pool's B200s, measured: tools/probe/dmma_probe.cu, profiles/R2_dmma_probe.txt


def fp64_pipe_block(leg, phases):
    """q >= 20 (tensor-pipe level kernels): flops of the two matrix products against the measured fp64 tensor peak.  The
    whole-pass HBM figure stays the headline bound; this block says how busy the pipe is that actually limits these
    kernels (DESIGN.md par. 4).  useful = 2 q^2 per product, child and pattern; issued = what the padded mma tiles execute."""
    q = leg.q
    flat = leg.flat
    n_int_children = int((flat['tip_row'][1:] < 0).sum())
    Lp = int(flat['multiplicity'].shape[0])
    ks = (q + 3) // 4
    mtiles = -(-Lp // 128) * 16                      # m-tiles of 8 patterns, whole 128-pattern tiles
    useful = {'postorder': 2.0 * q * q * n_int_children * Lp, 'preorder': 4.0 * q * q * n_int_children * Lp}
    issued = {'postorder': 512.0 * ks * 3 * n_int_children * mtiles, 'preorder': 1024.0 * ks * 3 * n_int_children * mtiles}
    out = {'peak_tflops': FP64_TENSOR_PEAK_TFLOPS, 'peak_source': 'measured (tools/probe/dmma_probe.cu; DFMA shares the pipe)'}
    for k in ('postorder', 'preorder'):
        sec = phases[k][0] / 1e3
        out[k] = {'useful_tflops': useful[k] / sec / 1e12, 'issued_tflops': issued[k] / sec / 1e12,
                  'frac_issued_of_peak': issued[k] / sec / 1e12 / FP64_TENSOR_PEAK_TFLOPS}
    return out


def roofline_block(leg, phases, pass_ms, workload):
    peak, peak_src = measured_peak()
    q = leg.q
    post_b, pre_b = algorithmic_bytes(leg.flat, q)
    dom = 'preorder' if phases['preorder'][0] >= phases['postorder'][0] else 'postorder'
    dom_bytes = pre_b if dom == 'preorder' else post_b
    dom_ms, dom_launches = phases[dom]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    mma = q > 8 and not leg.g.get('site_specific') and not os.environ.get('TTB_NO_MMA')
    extra = {}
    if mma:
        try:
            extra['fp64_tensor_pipe'] = fp64_pipe_block(leg, phases)
        except Exception as e:      # an extra, never the reason a bench line is lost
            extra['fp64_tensor_pipe'] = {'error': repr(e)}
    return dict(extra, **{
        'bound': 'hbm', 'kernel': '%s_level%s_kernel<%d> (%d level launches per pass)' % (
            'pre' if dom == 'preorder' else 'post', '_mma' if mma else '', q, dom_launches),
        'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
        'peak_note': 'peak is a measured COPY bandwidth (1:1 read:write); the preorder kernel streams 2:1 read:write and can sit at or slightly above it',
        'algorithmic_bytes_per_pass': int(dom_bytes), 'kernel_ms_per_pass': dom_ms,
        'traffic': ncu_traffic(workload, 'pre_level_kernel' if dom == 'preorder' else 'post_level_kernel'),
        'phases_ms': {k: v[0] for k, v in phases.items()}, 'phase_launches': {k: v[1] for k, v in phases.items()},
        'whole_pass': {'algorithmic_bytes': int(post_b + pre_b), 'bytes_per_update': (post_b + pre_b) / float(leg.updates_local),
                       'achieved_gbs': (post_b + pre_b) / (pass_ms / 1e3) / 1e9, 'frac': (post_b + pre_b) / (pass_ms / 1e3) / 1e9 / peak,
                       'survey_bytes_per_update': SURVEY_BYTES_PER_UPDATE.get(q),
                       'survey_frac': (SURVEY_BYTES_PER_UPDATE.get(q, 0) * leg.updates_local / (pass_ms / 1e3) / 1e9 / peak)
                       if q in SURVEY_BYTES_PER_UPDATE else None}})


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--scaling', default='auto', choices=['auto', 'strong', 'weak'],
                    help='N > 1: strong = one alignment, patterns sharded over the ranks (default); weak = one full-size alignment per rank')
    ap.add_argument('--cpu-patterns', type=int, default=0, help='patterns in the CPU sample (0 = auto)')
    ap.add_argument('--cpu-workers', type=int, default=0, help='processes of the CPU (reference) arm (0 = one per core)')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU port (drops `parity` and `cpu_baseline`)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-dense-e2e', action='store_true', help='skip the dense C-ABI e2e leg (N = 1)')
    ap.add_argument('--no-precision-study', action='store_true', help='skip the float-message-storage leg (N = 1)')
    ap.add_argument('--no-secondary', action='store_true', help='N > 1: skip the weak-scaling and north-star legs')
    ap.add_argument('--north-star', action='store_true', help='run the configs[4] leg at any N (default: only at N = 8)')
    ap.add_argument('--e2e-blocks', type=int, default=6, help='pattern blocks (engine handles / streams) of the dense C-ABI e2e leg')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    n_tips, L, alphabet, mean_bl, desc = WORKLOADS[args.workload]
    site_specific = alphabet == 'nuc_site_specific'
    mode = args.scaling
    if mode == 'auto':
        mode = 'strong' if (max(world, args.gpus) > 1 and not site_specific) else 'weak'
    if site_specific and mode == 'strong':
        mode = 'weak'        # the workload already IS one pattern shard per GPU of the 8-shard alignment
    args.scaling_resolved = mode

    if args.impl == 'reference':
        if rank != 0:
            return 0
        return run_reference_arm(args, real_stdout)

    import torch
    import torch.distributed as dist
    from treetime_b200.dist import SingleComm, TorchComm

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    stream = torch.cuda.Stream()          # a real (non-default) stream: events on it bracket the engine's work
    torch.cuda.set_stream(stream)
    k_e2e = max(3, min(args.steps, 10))

    def measure(leg, comm, sharded, with_parity, with_dense, workload_name):
        """All numbers of one leg; reductions over ranks inside.  Returns a dict (complete on rank 0)."""
        leg.prepare(comm)
        ms_local, launches, clocks, lh_global, lh_local, barrier = timed_resident(leg, args.steps, args.warmup, world, local_rank)
        leg.lh64 = lh_local
        phases = phase_profile(leg.eng)
        api = None
        if not args.no_e2e:
            api = timed_api_e2e(leg, k_e2e, barrier)
        dense = None
        if with_dense and not args.no_e2e and not args.no_dense_e2e:
            dense = dense_cabi_e2e(leg, k_e2e, args.e2e_blocks, local_rank, barrier)
        (ms_step, e2e_s, dense_s), (updates_total, h2d_tot, d2h_tot) = reduce_max_sum(
            world, [ms_local, api[0] if api else 0.0, dense[0] if dense else 0.0],
            [float(leg.updates_local), float(api[1]) if api else 0.0, float(api[2]) if api else 0.0])
        out = {'ms_per_step': ms_step, 'value': updates_total / (ms_step / 1e3), 'updates_per_step': updates_total,
               'patterns_rank0': int(leg.Lp), 'branches': int(leg.n_br), 'gpu_launches': int(launches), 'clocks': clocks,
               'log_lh': lh_global, 'device_bytes_rank0': leg.eng.device_bytes()}
        if rank == 0:
            out['roofline'] = roofline_block(leg, phases, ms_step, workload_name)
        if api:
            # the API's total is all-reduced by the TreeAnc's communicator when the patterns are sharded; the resident total
            # was all-reduced by the bench: global vs global (sharded) or local vs local (independent alignments)
            lh_api = api[3]
            lh_ref = lh_global if sharded else lh_local
            out['e2e'] = {'value': updates_total / e2e_s, 'unit': 'updates/s', 'ms_per_step': 1e3 * e2e_s,
                          'h2d_bytes_per_step': int(h2d_tot), 'd2h_bytes_per_step': int(d2h_tot), 'api': 'TreeAnc',
                          'rel_lh_diff_vs_resident_pass': abs(lh_api - lh_ref) / abs(lh_ref),
                          'alignment_differences_rank0': api[5], 'mutations_returned_rank0': api[4], 'host_binding': numa,
                          'what': 'per step through treetime_b200.TreeAnc (mirror of treetime.TreeAnc) with sparse_io: reload_alignment() '
                                  '[tip codes as reference row + differences, host -> device], infer_ancestral_sequences(marginal=True) '
                                  '[branch lengths + model up; tree.sequence_LH, total LH, N_diff down], sequence_differences() [every '
                                  'reconstructed sequence as root row + states differing from the parent, device -> host]; host wall clock, '
                                  'barrier + synchronize on both sides, max over ranks'}
        if dense:
            out['e2e_cabi_dense'] = {'value': updates_total / dense_s, 'unit': 'updates/s', 'ms_per_step': 1e3 * dense_s,
                                     'h2d_bytes_per_step': dense[1], 'd2h_bytes_per_step': dense[2], 'pattern_blocks': dense[3],
                                     'rel_lh_diff_vs_resident_pass': abs(dense[4] - lh_global) / abs(lh_global), 'pcie_probe_gbs': pcie_probe(),
                                     'what': 'C-ABI with dense host buffers: ttb_set_patterns (full tip-code matrix from pinned memory) / gtr / '
                                             'branch lengths, ttb_marginal, ttb_enqueue_fetch_site_lh + ttb_enqueue_fetch_all_seq_idx (every '
                                             'sequence dense) per pattern block on its own stream'}
        if with_parity and rank == 0 and not args.no_cpu_baseline:
            n_pat = args.cpu_patterns or int(max(48, min(leg.Lp, 20.0 * 2.0e6 / leg.n_br)))
            par, cpu_s, upd = parity_against_port(leg.eng, leg.flat, leg.g, n_pat)
            cached = par.pop('_cached')
            out['parity'] = par
            if world == 1 and leg.q <= 8 and not args.no_precision_study:
                out['precision_study'] = precision_study(leg, par, cached, args.steps, args.warmup, local_rank)
            out['cpu_baseline'] = {'value': upd / cpu_s, 'unit': 'updates/s', 'cores': 1, 'kind': 'port',
                                   'sample': 'one pass of oracle/flat_numpy.py over the first %d of %d patterns (same tree, same model); host has %d '
                                             'cores, the reference path is single-threaded numpy; `--impl reference` times the unmodified reference on '
                                             'all cores' % (par['patterns_compared'], leg.Lp, os.cpu_count() or 0)}
        return out

    sharded = world > 1 and mode == 'strong'
    primary_comm = TorchComm() if sharded else SingleComm()
    leg = Leg(args.workload, 1 if (sharded or world == 1) else 1 + rank, primary_comm, local_rank, stream)
    res = measure(leg, primary_comm, sharded, True, world == 1, args.workload)

    weak = None
    if world > 1 and mode == 'strong' and not args.no_secondary:
        # the weak form on the same engines: every rank the full alignment (the same columns everywhere -- per-GPU work is what counts)
        weak = measure(leg, SingleComm(), False, False, False, args.workload)
    del leg
    torch.cuda.empty_cache()

    ns = None
    if not site_specific and not args.no_secondary and (world == 8 or args.north_star):
        # BASELINE.json configs[4]: rank r holds shard r (3,750 uncompressed sites, its own per-site models) of the 30 kb alignment
        leg5 = Leg('cfg5', 1 + rank, SingleComm(), local_rank, stream)
        ns = measure(leg5, SingleComm(), False, True, False, 'cfg5')
        del leg5
        torch.cuda.empty_cache()

    if rank == 0:
        q = 5 if alphabet.startswith('nuc') else 20
        out = {
            'metric': METRIC, 'value': res['value'], 'unit': 'updates/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': mode, 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': '%s: %s' % (args.workload, desc), 'n_tips': n_tips, 'n_sites': L, 'n_states': q,
                       'patterns_rank0': res['patterns_rank0'], 'branches': res['branches'], 'updates_per_step': res['updates_per_step'],
                       'sharding': ('ONE alignment, compressed patterns sharded over %d ranks (dist.shard_bounds), tree replicated' % world) if sharded
                       else ('tree replicated, one full-size alignment per rank (%d rank%s)' % (world, 's' if world > 1 else '')),
                       'l2': 'working set (%.1f GB on rank 0) vs 126 MB L2: no flush needed' % (res['device_bytes_rank0'] / 1e9),
                       'device_bytes_rank0': res['device_bytes_rank0']},
            'clocks': res['clocks'], 'gpu_launches': res['gpu_launches'], 'log_lh': res['log_lh'], 'roofline': res['roofline']}
        for k in ('parity', 'cpu_baseline', 'e2e', 'e2e_cabi_dense', 'precision_study'):
            if k in res:
                out[k] = res[k]
        if 'parity' in res:
            out['log_lh_rel_err'] = res['parity']['log_lh_rel_err']
            out['max_profile_abs_err'] = res['parity']['max_profile_abs_err']
            out['argmax_mismatch_off_ties'] = res['parity']['argmax_mismatch_off_ties']
        if weak is not None:
            out['weak_scaling'] = {k: weak[k] for k in ('value', 'ms_per_step', 'updates_per_step', 'patterns_rank0', 'clocks', 'e2e') if k in weak}
            out['weak_scaling']['what'] = 'same run, every rank a full-size configs[2] alignment (per-GPU work fixed)'
        if ns is not None:
            blk = {'workload': 'cfg5: ' + WORKLOADS['cfg5'][4], 'n_gpus': world, 'shards_run': world, 'shards_total': 8}
            blk.update({k: ns[k] for k in ('value', 'ms_per_step', 'updates_per_step', 'patterns_rank0', 'branches', 'clocks', 'gpu_launches',
                                           'roofline', 'parity', 'e2e', 'log_lh') if k in ns})
            blk['whole_pass_frac_of_hbm_roofline'] = ns['roofline']['whole_pass']['frac']
            out['north_star_config'] = blk
        _emit(real_stdout, out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
