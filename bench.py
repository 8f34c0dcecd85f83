#!/usr/bin/env python
"""bench.py -- marginal ancestral reconstruction throughput (branch x pattern updates/s).

    python bench.py [--gpus N --steps K --warmup W] [--workload cfg3|cfg2|cfg1|cfg4|tiny]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # CPU arm: the oracle port of the reference's numpy path

One step = one full `infer_ancestral_sequences(marginal=True)` pass (batched expQt, level-ordered
postorder, root, level-ordered preorder, reductions) over one synthetic alignment shard.
N > 1: one process per GPU, the tree/model replicated, every rank owns its own block of
alignment columns (weak scaling: the per-GPU shard is fixed), the only collective is the
all-reduce of {total log-LH, N_diff}.

Prints ONE JSON line (rank 0).  `value` = updates/s with inputs resident in HBM, timed with CUDA
events around the K steps (max over ranks); `e2e` = the same metric through the C-ABI with HOST
buffers: per step the tip codes / branch lengths / model are copied host->device from pinned
memory and the per-pattern LH, the totals and every reconstructed sequence are copied back.
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
    'cfg4': (5000, 1000, 'aa_nogap', 1e-2, '5k tips x 1,000-site amino-acid alignment, 20-state model (BASELINE.json configs[3])'),
    'cfg5': (100000, 3750, 'nuc_site_specific', 3.3e-5, '100k tips x 30 kb nucleotide with site-specific GTR, one of 8 pattern shards '
                                                       '(3,750 uncompressed sites per GPU) (BASELINE.json configs[4])'),
    'tiny': (64, 500, 'nuc', 1e-2, 'smoke-sized'),
}
SURVEY_BYTES_PER_UPDATE = {5: 165.5, 4: 133.5, 20: 645.5, 22: 709.5}    # 4 q s + 0.5 s + 1.5


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and so its first-touch pinned buffers) to the CPUs next to its GPU: with
    several ranks the end-to-end leg is otherwise limited by cross-socket host traffic, not by PCIe."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), 'pci_domain_id', 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), 'pci_device_id', 0)
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0' % (dom, bus, dev)
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


def make_workload(name, seed):
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    n_tips, L, alphabet, mean_bl, _ = WORKLOADS[name]
    compress = True
    if alphabet == 'nuc':
        gtr = GTR.custom(pi=np.array([0.3, 0.2, 0.2, 0.29, 0.01]), W=np.ones((5, 5)), alphabet='nuc')
    elif alphabet == 'nuc_site_specific':
        from treetime_b200.gtr import GTRSiteSpecific
        gtr = GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(1000 + seed))
        compress = False          # treeanc.py:186-187: no pattern compression with site-specific models
    else:
        gtr = GTR.random(alphabet=alphabet, rng=np.random.default_rng(1234))
    tree = synth.random_tree(n_tips, seed=1, mean_bl=mean_bl)           # same tree on every rank
    topo, flat, g = synth.make_flat_problem(tree, gtr, L, seed, compress=compress)   # rank-specific columns
    return topo, flat, g


def algorithmic_bytes(flat, q):
    """Bytes the level kernels must move per pass under THIS design:
    postorder: read S of internal children + 1-byte codes of tip children, write S per internal
    node (the log-prefactors are summed per block run, not stored per node); preorder: read parent profile once per parent with an internal child, per
    internal child read S, write profile, read+write the 1-byte state."""
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
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


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


def run_cpu(flat, g, n_patterns, repeats):
    """Time the oracle port (oracle/flat_numpy.py: the reference's per-node numpy calls)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import flat_numpy as O
    s, n, g = cpu_sample(flat, g, n_patterns)
    n_br = flat['parent'].shape[0] - 1
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = O.marginal(s, g)
        times.append(time.perf_counter() - t0)
    return n_br * n, times, res.total_LH


_POOL_STATE = {}


def _cpu_worker_init(flat, g):
    """Worker of the multi-process CPU arm: one pattern slice per process, numpy pinned to one thread."""
    try:
        import threadpoolctl
        _POOL_STATE['limit'] = threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    _POOL_STATE['flat'], _POOL_STATE['g'] = flat, g


def _cpu_worker_run(bounds):
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


def run_cpu_parallel(flat, g, per_worker, workers, repeats):
    """The reference path is single-threaded numpy; patterns are independent, so the strongest CPU arm this host
    offers is one process per core, each running the oracle port on its own slice of the pattern axis."""
    import multiprocessing as mp
    Lp = flat['multiplicity'].shape[0]
    per_worker = max(1, min(per_worker, Lp // workers))
    bounds = [(k * per_worker, (k + 1) * per_worker) for k in range(workers)]
    n_br = flat['parent'].shape[0] - 1
    times = []
    with mp.get_context('fork').Pool(workers, initializer=_cpu_worker_init, initargs=(flat, g)) as pool:
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker_run, bounds, chunksize=1)
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


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--cpu-patterns', type=int, default=0, help='patterns in the CPU baseline sample (0 = auto)')
    ap.add_argument('--cpu-workers', type=int, default=0, help='processes of the CPU (reference) arm (0 = one per core)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-blocks', type=int, default=6, help='pattern blocks (engine handles / streams) of the e2e leg')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    n_tips, L, alphabet, mean_bl, desc = WORKLOADS[args.workload]

    # ------------------------------------------------------------------ CPU (reference) arm
    if args.impl == 'reference':
        if rank != 0:
            return 0
        topo, flat, g = make_workload(args.workload, seed=1)
        q = g['Pi'].shape[0]
        n_br = flat['parent'].shape[0] - 1
        total_steps = args.steps + args.warmup
        n_pat = args.cpu_patterns or int(max(64, min(flat['multiplicity'].shape[0], 120.0 * 2.0e6 / (n_br * max(1, total_steps)))))
        # one process per host core, bounded by memory: the port keeps ~6 (n_nodes, patterns, q) fp64 arrays per slice
        workers = max(1, min(os.cpu_count() or 1, args.cpu_workers or 64))
        per_worker = max(64, min(n_pat, flat['multiplicity'].shape[0] // workers))
        try:
            import psutil
            per_pattern = 6.0 * flat['parent'].shape[0] * q * 8          # bytes one pattern costs a worker
            budget = 0.5 * psutil.virtual_memory().available
            per_worker = int(min(per_worker, max(64, budget // (workers * per_pattern))))
            workers = int(max(1, min(workers, budget // (per_worker * per_pattern))))
        except Exception:
            pass
        if workers > 1:
            updates, times, per_worker = run_cpu_parallel(flat, g, per_worker, workers, total_steps)
            sample = ('%d processes (one per core), each one pass over its own %d of %d compressed patterns of the same '
                      'tree/alignment per step (cost is linear in patterns)' % (workers, per_worker, flat['multiplicity'].shape[0]))
        else:
            updates, times, _ = run_cpu(flat, g, n_pat, total_steps)
            sample = 'first %d of %d compressed patterns of the same tree/alignment per step (cost is linear in patterns)' % (
                min(n_pat, flat['multiplicity'].shape[0]), flat['multiplicity'].shape[0])
        timed = times[args.warmup:]
        ms = 1e3 * float(np.mean(timed))
        val = updates / (ms / 1e3)
        _emit(real_stdout, {
            'impl': 'reference', 'metric': 'marginal ancestral reconstruction branch x pattern updates/s', 'value': val,
            'unit': 'updates/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': '%s: %s' % (args.workload, desc), 'n_tips': n_tips, 'n_sites': L, 'n_states': q},
            'cpu_baseline': {'value': val, 'unit': 'updates/s', 'cores': workers, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': 'oracle/flat_numpy.py: flat-array port of the reference numpy path (bit-identical to the '
                    'reference on the build container).  The reference itself is single-threaded; patterns are '
                    'independent, so this arm runs one process per host core on disjoint pattern slices',
        })
        return 0

    # ------------------------------------------------------------------ GPU arm
    import torch
    import torch.distributed as dist
    from treetime_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    topo, flat, g = make_workload(args.workload, seed=1 + rank)
    q = g['Pi'].shape[0]
    n_nodes = flat['parent'].shape[0]
    n_br = n_nodes - 1
    Lp = flat['multiplicity'].shape[0]
    updates_local = n_br * Lp

    stream = torch.cuda.Stream()          # a real (non-default) stream: events on it bracket the engine's work
    torch.cuda.set_stream(stream)
    eng = Engine(q, device=local_rank)
    eng.set_stream(stream.cuda_stream)
    eng.set_tree(flat['parent'], flat['child_ptr'], flat['child_idx'], flat['tip_row'])
    # pinned host staging for the e2e leg
    codes_pin = torch.empty(flat['tip_codes'].shape, dtype=torch.uint8, pin_memory=True)
    codes_pin.numpy()[...] = flat['tip_codes']
    eng.set_patterns(codes_pin.numpy(), flat['code_profiles'], flat['multiplicity'])
    eng.set_gtr(g)
    eng.set_branch_lengths(flat['t'])

    class _Dev(object):
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {'shape': (2,), 'typestr': '<f8', 'data': (ptr, False), 'version': 2}

    def step():
        eng.marginal()
        if world > 1:
            dist.all_reduce(res_t)

    eng.marginal()
    eng.sync()
    res_t = torch.as_tensor(_Dev(eng.results_device_ptr()), device='cuda') if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    total_lh_local, _ = eng.results() if world == 1 else (float(res_t[0].item()), 0)

    # per-phase device times (un-graphed pass, CUDA events between phases)
    phases = eng.profile_marginal()
    for _ in range(2):
        ph = eng.profile_marginal()
        phases = {k: (min(phases[k][0], ph[k][0]), ph[k][1]) for k in ph}

    # ---- e2e: host buffers in, host results out, every step.
    # The pattern axis is cut into E2E_BLOCKS column blocks, each with its own engine handle and
    # stream (one handle = one pattern shard is the C-ABI's unit anyway), so the H2D copy of block
    # k+1, the pass over block k and the D2H copy of block k-1 overlap on the two copy engines.
    e2e = None
    if not args.no_e2e:
        n_int = int((flat['tip_row'] < 0).sum())
        nblk = max(1, min(args.e2e_blocks, Lp // 1024))
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

        def e2e_step():
            # software pipeline over the blocks: uploads are chained (block k+1's copy starts when block k's
            # has landed) and so are the passes; otherwise the DMA engine and the SMs time-slice all
            # blocks and nothing overlaps
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
            tot = 0.0
            for e, *_ in shards:
                t_, _ = e.results()          # waits for that block's stream (incl. its D2H copies)
                tot += t_
            if world > 1:
                tt_ = torch.tensor([tot], device='cuda', dtype=torch.float64)
                dist.all_reduce(tt_)
                tot = float(tt_.item())
            return tot

        tot_e2e = e2e_step()
        e2e_step()
        barrier()
        k_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            tot_e2e = e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / k_e2e
        h2d = int(flat['tip_codes'].nbytes + nblk * (flat['code_profiles'].nbytes + flat['t'].nbytes + 8 * (2 * q * q + 2 * q + 1))
                  + flat['multiplicity'].nbytes)
        d2h = int(n_int * Lp + 8 * Lp + 16 * nblk)
        # sanity: the blocked e2e result equals the resident pass
        lh_check = abs(tot_e2e - (total_lh_local if world == 1 else float(res_t[0].item()))) / abs(tot_e2e)
        # PCIe probe for context (pinned 256 MB each way)
        pb = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
        db = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        pe0, pe1, pe2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        db.copy_(pb, non_blocking=True); torch.cuda.synchronize()
        pe0.record(); db.copy_(pb, non_blocking=True); pe1.record(); pb.copy_(db, non_blocking=True); pe2.record()
        torch.cuda.synchronize()
        pcie = (0.268435456 / (pe0.elapsed_time(pe1) / 1e3), 0.268435456 / (pe1.elapsed_time(pe2) / 1e3))
        for sh in shards:
            sh[0].close()
        del shards
        torch.cuda.empty_cache()
        # ---- sparse host interface: the alignment goes in as (reference row + differences), the
        # sequences come back as (root row + states that differ from the parent).  Same information,
        # ~100x fewer PCIe bytes; one handle, no blocking needed.
        from treetime_b200.sparse import sparse_from_dense
        ref_c, e_row, e_pos, e_code = sparse_from_dense(flat['tip_codes'])
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()  # noqa: E731
        ref_c, e_row, e_pos, e_code = pin(ref_c), pin(e_row), pin(e_pos), pin(e_code)
        max_mut = max(1 << 16, 8 * n_nodes)
        m_node = torch.empty(max_mut, dtype=torch.int32, pin_memory=True).numpy()
        m_pos = torch.empty(max_mut, dtype=torch.int32, pin_memory=True).numpy()
        m_state = torch.empty(max_mut, dtype=torch.uint8, pin_memory=True).numpy()
        m_root = torch.empty(Lp, dtype=torch.uint8, pin_memory=True).numpy()
        lh_pin = torch.empty(Lp, dtype=torch.float64, pin_memory=True).numpy()
        import ctypes
        from treetime_b200 import _lib as L_
        from treetime_b200.engine import _ip, _up

        def sparse_step():
            eng.set_patterns_sparse(ref_c, e_row, e_pos, e_code, flat['code_profiles'], flat['multiplicity'])
            eng.set_gtr(g)
            eng.set_branch_lengths(flat['t'])
            eng.marginal()
            if world > 1:
                dist.all_reduce(res_t)
            eng.enqueue_site_lh(lh_pin)
            n_ = ctypes.c_int64()
            L_.check(eng.lib.ttb_fetch_mutations(eng.h, _up(m_root), max_mut, _ip(m_node), _ip(m_pos), _up(m_state), ctypes.byref(n_)))
            tot_, _ = eng.results()
            return tot_, int(n_.value)

        sparse_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            tot_sp, n_mut = sparse_step()
        barrier()
        sp_s = (time.perf_counter() - t0) / k_e2e
        sp_h2d = int(ref_c.nbytes + e_row.nbytes + e_pos.nbytes + e_code.nbytes + flat['code_profiles'].nbytes + flat['multiplicity'].nbytes
                     + flat['t'].nbytes + 8 * (2 * q * q + 2 * q + 1))
        sp_d2h = int(Lp + n_mut * 9 + 8 * Lp + 24)
        sp_check = abs(tot_sp - (total_lh_local if world == 1 else float(res_t[0].item()))) / abs(tot_sp)
        e2e = (e2e_s, h2d, d2h, nblk, lh_check, pcie, sp_s, sp_h2d, sp_d2h, sp_check, n_mut, int(e_row.shape[0]))

    # ---- reduce over ranks: max time, summed work
    ms_step = ms_total / args.steps
    if world > 1:
        tmax = torch.tensor([ms_step, e2e[0] if e2e else 0.0, e2e[6] if e2e else 0.0], device='cuda', dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = torch.tensor([float(updates_local)], device='cuda', dtype=torch.float64)
        dist.all_reduce(tsum)
        ms_step = float(tmax[0].item())
        e2e_time = float(tmax[1].item())
        sp_time = float(tmax[2].item())
        updates_total = float(tsum[0].item())
    else:
        e2e_time = e2e[0] if e2e else 0.0
        sp_time = e2e[6] if e2e else 0.0
        updates_total = float(updates_local)

    if rank == 0:
        value = updates_total / (ms_step / 1e3)
        peak, peak_src = measured_peak()
        post_b, pre_b = algorithmic_bytes(flat, q)
        dom = 'preorder' if phases['preorder'][0] >= phases['postorder'][0] else 'postorder'
        dom_bytes = pre_b if dom == 'preorder' else post_b
        dom_ms, dom_launches = phases[dom]
        achieved = dom_bytes / (dom_ms / 1e3) / 1e9
        # whole-pass figure: the timed graph replays (ms_step), not the sum of the separately timed phases
        pass_ms = ms_step
        out = {
            'metric': 'marginal ancestral reconstruction branch x pattern updates/s', 'value': value, 'unit': 'updates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': '%s: %s' % (args.workload, desc), 'n_tips': n_tips, 'n_sites_per_gpu': L, 'n_states': q,
                       'patterns_per_gpu': int(Lp), 'branches': int(n_br), 'updates_per_step': updates_total,
                       'sharding': 'pattern blocks, tree replicated (%d rank%s)' % (world, 's' if world > 1 else ''),
                       'l2': 'working set (%.1f GB/GPU) far larger than the 126 MB L2; no flush needed' % (eng.device_bytes() / 1e9),
                       'device_bytes_per_gpu': eng.device_bytes()},
            'clocks': clocks,
            'gpu_launches': int(launches),
            'log_lh_rank0': total_lh_local,
            'roofline': {
                'bound': 'hbm', 'kernel': '%s_level_kernel<%d> (%d level launches per pass)' % ('pre' if dom == 'preorder' else 'post', q, dom_launches),
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
                'peak_note': 'peak is a measured COPY bandwidth (1:1 read:write); the preorder kernel streams 2:1 read:write and can sit at or slightly above it',
                'algorithmic_bytes_per_pass': int(dom_bytes), 'kernel_ms_per_pass': dom_ms,
                'traffic': ncu_traffic(args.workload, 'pre_level_kernel' if dom == 'preorder' else 'post_level_kernel'),
                'phases_ms': {k: v[0] for k, v in phases.items()}, 'phase_launches': {k: v[1] for k, v in phases.items()},
                'whole_pass': {'algorithmic_bytes': int(post_b + pre_b), 'bytes_per_update': (post_b + pre_b) / float(updates_local),
                               'achieved_gbs': (post_b + pre_b) / (pass_ms / 1e3) / 1e9,
                               'frac': (post_b + pre_b) / (pass_ms / 1e3) / 1e9 / peak,
                               'survey_bytes_per_update': SURVEY_BYTES_PER_UPDATE.get(q),
                               'survey_frac': (SURVEY_BYTES_PER_UPDATE.get(q, 0) * updates_local / (pass_ms / 1e3) / 1e9 / peak)
                               if q in SURVEY_BYTES_PER_UPDATE else None},
            },
        }
        if e2e:
            out['e2e_sparse_io'] = {
                'value': updates_total / sp_time, 'unit': 'updates/s', 'ms_per_step': 1e3 * sp_time,
                'h2d_bytes_per_step': e2e[7], 'd2h_bytes_per_step': e2e[8], 'rel_lh_diff_vs_resident_pass': e2e[9],
                'alignment_differences': e2e[11], 'mutations_returned': e2e[10],
                'what': 'same pass through ttb_set_patterns_sparse (reference row + differences, like TreeTime\'s VCF '
                        'alignments) and ttb_fetch_mutations (root row + states differing from the parent) + per-pattern LH'}
            out['e2e'] = {'value': updates_total / e2e_time, 'unit': 'updates/s', 'ms_per_step': 1e3 * e2e_time,
                          'h2d_bytes_per_step': e2e[1], 'd2h_bytes_per_step': e2e[2],
                          'host_binding': numa, 'pattern_blocks': e2e[3], 'rel_lh_diff_vs_resident_pass': e2e[4],
                          'pcie_probe_gbs': {'h2d': e2e[5][0], 'd2h': e2e[5][1]},
                          'what': 'per step and per pattern block: ttb_set_patterns/gtr/branch_lengths from pinned host '
                                  'memory, ttb_marginal, ttb_enqueue_fetch_site_lh + ttb_enqueue_fetch_all_seq_idx into '
                                  'pinned host memory, ttb_results; blocks run on their own streams so copies overlap compute'}
        if not args.no_cpu_baseline and world >= 1:
            n_pat = args.cpu_patterns or int(max(64, min(Lp, 20.0 * 2.0e6 / n_br)))
            upd, times, cpu_lh = run_cpu(flat, g, n_pat, 1)
            out['cpu_baseline'] = {'value': upd / times[0], 'unit': 'updates/s', 'cores': 1, 'kind': 'port',
                                   'sample': 'one pass over the first %d of %d patterns (same tree, same model); '
                                             'host has %d cores, the reference path is single-threaded numpy'
                                             % (min(n_pat, Lp), Lp, os.cpu_count() or 0)}
        _emit(real_stdout, out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
