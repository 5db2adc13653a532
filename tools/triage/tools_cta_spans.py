"""perf triage: wall-clock spans (ns, %globaltimer) of every CTA of the forward kernel:
entry -> prologue done -> role loops done -> exit; run for the full kernel and the all-ablated skeleton"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench_utils import surface_cloud
from spconv_b200.core import ConvAlgo
from spconv_b200.pytorch import ops

def _dbg(debug=0, trace=None, ctas=0):
    """perf-triage hooks go through the explicit C-ABI call (spx_debug_configure), not the environment"""
    from spconv_b200 import _cabi as _c
    _c.check(_c.load().spx_debug_configure(-1, int(ctas), int(debug), None if trace is None else trace.data_ptr(),
                                            0 if trace is None else trace.numel() * trace.element_size()), "debug_configure")
dev = torch.device("cuda:0")
shape=[41,1600,1408]; C=K=64
rng=np.random.default_rng(50051)
inds=torch.from_numpy(surface_cloud(rng,shape,100000)).to(dev)
x=torch.randn(100000,C,device=dev).half(); w=(torch.randn(K,3,3,3,C,device=dev)*0.05).half()
res=ops.get_indice_pairs_implicit_gemm(inds,1,shape,ConvAlgo.MaskImplicitGemm,[3]*3,[1]*3,[1]*3,[1]*3,[0]*3,True,False,is_train=True)
_,_,pf,pb,mf,mb,sf,sb,masks=res
fwd=lambda: ops.implicit_gemm(x,w,pf,mf,sf,100000,masks,True,True)
for _ in range(3): fwd()
for dbg in ("0","15","1","2","4"):
    ts=torch.zeros((8,2048),dtype=torch.int64,device=dev)
    _dbg(int(dbg), ts)
    torch.cuda._sleep(400000)
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record(); fwd(); b.record(); torch.cuda.synchronize()
    _dbg(0, None)
    t=ts.cpu().numpy()[6:8].reshape(-1)[:1024*4].reshape(-1,4)
    t=t[t[:,0]>0]
    t0=t[:,0].min()
    r=(t-t0)/1000.0
    print(f"debug={dbg}: event time {a.elapsed_time(b)*1000:.1f} us, CTAs {len(r)}")
    print(f"   entry      : min {r[:,0].min():6.2f}  median {np.median(r[:,0]):6.2f}  max {r[:,0].max():6.2f} us")
    print(f"   prologue   : median {np.median(r[:,1]-r[:,0]):6.2f}  max {(r[:,1]-r[:,0]).max():6.2f} us")
    print(f"   role loops : min {(r[:,2]-r[:,1]).min():6.2f}  median {np.median(r[:,2]-r[:,1]):6.2f}  max {(r[:,2]-r[:,1]).max():6.2f} us")
    print(f"   exit       : median {np.median(r[:,3]):6.2f}  max {r[:,3].max():6.2f} us (last CTA done)")
    d=r[:,2]-r[:,1]
    print("   loop-duration histogram (us):", np.histogram(d,bins=8)[0].tolist(), [round(v,1) for v in np.histogram(d,bins=8)[1].tolist()])
