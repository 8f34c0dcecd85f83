"""Host-side profile (cProfile) of the product-API e2e step of bench.py at cfg3: where the 11 ms beyond the resident pass go."""
import cProfile, pstats, sys, time
sys.path.insert(0, '.')
import torch
import bench
from treetime_b200.dist import SingleComm
leg = bench.Leg('cfg3', 1, SingleComm(), 0, torch.cuda.Stream())
leg.prepare(SingleComm())
tt = leg.tt


def step():
    tt.reload_alignment()
    tt.infer_ancestral_sequences(marginal=True)
    return tt.sequence_differences(gather=False)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
print('e2e step: %.2f ms' % ((time.perf_counter() - t0) * 100))
for name, f in (('reload+infer', lambda: (tt.reload_alignment(), tt.infer_ancestral_sequences(marginal=True))), ('sequence_differences', lambda: tt.sequence_differences(gather=False))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        f()
    torch.cuda.synchronize()
    print('%s: %.2f ms' % (name, (time.perf_counter() - t0) * 100))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
