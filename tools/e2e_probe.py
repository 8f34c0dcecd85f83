"""Timeline probe of the blocked e2e leg: CUDA events per pattern block (H2D end, pass end, D2H end)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, bench
from treetime_b200.engine import Engine
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 4
topo, flat, g = bench.make_workload('cfg3', 1)
q=5; Lp=flat['multiplicity'].shape[0]; n_int=int((flat['tip_row']<0).sum())
bounds=[(Lp*i)//nblk for i in range(nblk+1)]
shards=[]
for i in range(nblk):
    lo,hi=bounds[i],bounds[i+1]
    st=torch.cuda.Stream()
    e=Engine(q); e.set_stream(st.cuda_stream); e.set_tree(flat['parent'],flat['child_ptr'],flat['child_idx'],flat['tip_row'])
    cp=torch.empty((flat['tip_codes'].shape[0],hi-lo),dtype=torch.uint8,pin_memory=True); cp.numpy()[...]=flat['tip_codes'][:,lo:hi]
    sp=torch.empty((n_int,hi-lo),dtype=torch.uint8,pin_memory=True); lp=torch.empty(hi-lo,dtype=torch.float64,pin_memory=True)
    shards.append((e,cp.numpy(),sp.numpy(),lp.numpy(),np.ascontiguousarray(flat['multiplicity'][lo:hi]),st,cp,sp,lp))
def step(trace=False):
    ev=[]
    t0e=torch.cuda.Event(enable_timing=True); t0e.record(torch.cuda.default_stream()); 
    torch.cuda.synchronize(); t0=time.perf_counter()
    base=torch.cuda.Event(enable_timing=True); base.record(shards[0][5])
    for e,cp,sp,lp,m,st,*_ in shards:
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e2=torch.cuda.Event(enable_timing=True); e3=torch.cuda.Event(enable_timing=True)
        e0.record(st); e.set_patterns(cp,flat['code_profiles'],m,validate=False); e.set_gtr(g); e.set_branch_lengths(flat['t']); e1.record(st)
        e.marginal(); e2.record(st); e.enqueue_site_lh(lp); e.enqueue_all_seq_idx(sp); e3.record(st)
        ev.append((e0,e1,e2,e3))
    for e,*_ in shards: e.results()
    torch.cuda.synchronize(); z=time.perf_counter()
    if trace:
        for i,(e0,e1,e2,e3) in enumerate(ev):
            print('block %d: start %.2f  h2d_end %.2f  pass_end %.2f  d2h_end %.2f' % (i, base.elapsed_time(e0), base.elapsed_time(e1), base.elapsed_time(e2), base.elapsed_time(e3)))
        print('total %.2f ms' % (1e3*(z-t0)))
for _ in range(3): step()
step(True)
