import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, bench
from treetime_b200.engine import Engine
topo, flat, g = bench.make_workload('cfg3', 1)
q=5; Lp=flat['multiplicity'].shape[0]; n_int=int((flat['tip_row']<0).sum())
nblk=4; bounds=[(Lp*i)//nblk for i in range(nblk+1)]
shards=[]
for i in range(nblk):
    lo,hi=bounds[i],bounds[i+1]
    e=Engine(q); e.set_tree(flat['parent'],flat['child_ptr'],flat['child_idx'],flat['tip_row'])
    cp=torch.empty((flat['tip_codes'].shape[0],hi-lo),dtype=torch.uint8,pin_memory=True); cp.numpy()[...]=flat['tip_codes'][:,lo:hi]
    sp=torch.empty((n_int,hi-lo),dtype=torch.uint8,pin_memory=True); lp=torch.empty(hi-lo,dtype=torch.float64,pin_memory=True)
    shards.append((e,cp.numpy(),sp.numpy(),lp.numpy(),np.ascontiguousarray(flat['multiplicity'][lo:hi]),cp,sp,lp))
def step(trace=False):
    T=[]; t0=time.perf_counter()
    for e,cp,sp,lp,m,*_ in shards:
        a=time.perf_counter(); e.set_patterns(cp,flat['code_profiles'],m,validate=False); b=time.perf_counter(); e.set_gtr(g); c=time.perf_counter(); e.set_branch_lengths(flat['t']); d=time.perf_counter(); e.marginal(); f=time.perf_counter(); e.enqueue_site_lh(lp); e.enqueue_all_seq_idx(sp); h=time.perf_counter()
        T.append([round(1e3*x,2) for x in (b-a,c-b,d-c,f-d,h-f)])
    r=time.perf_counter()
    for e,*_ in shards: e.results()
    z=time.perf_counter()
    if trace: print('per block [set_patterns,set_gtr,set_t,marginal,enqueue fetch] ms:',T,' wait %.2f total %.2f'%(1e3*(z-r),1e3*(z-t0)))
for _ in range(3): step()
torch.cuda.synchronize(); step(True); step(True)
