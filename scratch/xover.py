import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O
from ccspecs import model_variants, controller_variants
from cchelpers import conv
from chain_control_b200 import ops
dev='cuda'; T_=lambda a: torch.tensor(np.asarray(a,np.float32),device=dev)
ms_=model_variants()['compact_s50_d0']; cs_=controller_variants()['compact_s5_d0']
spec=conv(ms_); cs=conv(cs_); theta=T_(O.init_params(ms_,1,stable=True)); thc=T_(O.init_params(cs_,2))
def timeit(fn,n=3):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for B in (256, 512, 768, 1024, 1184, 1536, 2048, 4096):
    Tr=1001; refs=torch.rand(B,Tr,1,device=dev)*2-1; ybc=torch.randn(B,Tr,1,device=dev)/(B*Tr); u=refs[:, :1000].contiguous(); yb=torch.randn(B,1001,1,device=dev)
    row=[]
    for fam,(tc,mx) in {'latency':('1',str(1<<40)),'tcgen05':('1','0')}.items():
        os.environ['CC_B200_TC']=tc; os.environ['CC_B200_WPT_MAXB']=mx
        c0=timeit(lambda: ops.closedloop_fwd(cs,spec,thc,theta,refs,want_tape=True))
        yh,_,tape=ops.closedloop_fwd(cs,spec,thc,theta,refs,want_tape=True)
        cb=timeit(lambda: ops.closedloop_bwd(cs,spec,thc,theta,refs,tape,ybc))
        f0=timeit(lambda: ops.rollout_fwd(spec,theta,u,want_tape=True))
        y,_,tp=ops.rollout_fwd(spec,theta,u,want_tape=True)
        ob=timeit(lambda: ops.rollout_bwd(spec,theta,u,tp,yb))
        row.append('%s closed %.2f+%.2f=%.2f open %.2f+%.2f=%.2f' % (fam,c0,cb,c0+cb,f0,ob,f0+ob))
    print('B=%5d | %s | %s' % (B,row[0],row[1]),flush=True)
