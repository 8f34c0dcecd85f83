"""Pattern sharding across GPUs and the (tiny) collectives it needs.

Alignment columns are conditionally independent given tree + model, so each rank
owns a contiguous block of the compressed-pattern axis, the tree / branch lengths
/ GTR are replicated, and the only data that ever crosses NVLink is
  * the total log-likelihood and N_diff            (2 doubles per pass),
  * per Brent iteration one vector of n_branches partial objective values,
  * the q^2 + q substitution statistics when a GTR is inferred.
One process per GPU; the collectives go through torch.distributed (NCCL on GPU,
gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_patterns, rank, world_size):
    """Contiguous, balanced-by-count block [lo, hi) of the pattern axis for `rank`."""
    base, rem = divmod(int(n_patterns), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SingleComm(object):
    """world_size = 1: every collective is the identity."""
    rank, world_size = 0, 1

    def allreduce_sum(self, x):
        return np.asarray(x, dtype=np.float64)

    def allgather(self, x, axis=0, sizes=None):
        return np.asarray(x)

    def barrier(self):
        pass


class TorchComm(object):
    """Collectives over an initialised torch.distributed process group."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised')
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        backend = dist.get_backend(group)
        if device is None:
            device = ('cuda:%d' % torch.cuda.current_device()) if backend == 'nccl' else 'cpu'
        self.device = torch.device(device)

    def allreduce_sum(self, x):
        t = self.torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allgather(self, x, axis=0, sizes=None):
        """Concatenate per-rank arrays (possibly of different length along `axis`).  Tensor collectives, no pickling:
        the per-rank lengths (skipped when the caller knows them -- pattern shards follow shard_bounds -- and passes
        `sizes`), then the data padded to the longest shard into ONE output tensor, so a gather costs one collective, one
        host->device and one device->host copy whatever the world size (it runs every pass for tree.sequence_LH)."""
        torch = self.torch
        x = np.ascontiguousarray(np.moveaxis(np.asarray(x), axis, 0))
        W = self.world_size
        if sizes is None:
            n = torch.tensor([x.shape[0]], dtype=torch.int64, device=self.device)
            got = self._gather_flat(n, W)
            sizes = [int(v) for v in got.cpu().tolist()]
        else:
            sizes = [int(v) for v in sizes]
            assert len(sizes) == W and sizes[self.rank] == x.shape[0], 'allgather: sizes do not match this rank\'s array'
        width = max(sizes)
        t = torch.zeros((width,) + x.shape[1:], dtype=torch.from_numpy(x[:0]).dtype, device=self.device)
        if x.shape[0]:
            t[:x.shape[0]] = torch.from_numpy(x).to(self.device)
        host = self._gather_flat(t, W).cpu().numpy().reshape((W, width) + x.shape[1:])
        out = np.concatenate([host[r, :k] for r, k in enumerate(sizes)], axis=0)
        return np.moveaxis(out, 0, axis)

    def _gather_flat(self, t, W):
        """all_gather of equally shaped tensors into one tensor of W times the leading extent."""
        out = self.torch.empty((W * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        try:
            self.dist.all_gather_into_tensor(out, t, group=self.group)
        except (RuntimeError, NotImplementedError, AttributeError):      # a backend without the fused form
            parts = [self.torch.empty_like(t) for _ in range(W)]
            self.dist.all_gather(parts, t, group=self.group)
            out = self.torch.cat(parts, dim=0)
        return out

    def allreduce_sum_device(self, engine):
        """In-place sum over the ranks of n doubles at a device address (NCCL only): a callable (ptr, n) for
        Engine.brent_minimize, or None when the group cannot reduce device memory (gloo).  The engine runs on its own
        stream: it is drained before the collective, and the collective before the engine continues."""
        if self.dist.get_backend(self.group) != 'nccl':
            return None
        torch = self.torch

        class _View(object):
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f8', 'data': (ptr, False), 'version': 2}

        def reduce(ptr, n):
            engine.sync()
            t = torch.as_tensor(_View(ptr, n), device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            torch.cuda.current_stream(self.device).synchronize()
        return reduce

    def barrier(self):
        self.dist.barrier(group=self.group)


def default_comm():
    """TorchComm when torch.distributed is initialised with world_size > 1, else SingleComm."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchComm()
    except ImportError:
        pass
    return SingleComm()
