"""perf triage: per-stage clock64 timeline of CTA 0 of the forward kernel (SPX_TC_TRACE)"""
import os, sys, json
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
for _ in range(3): ops.implicit_gemm(x,w,pf,mf,sf,100000,masks,True,True)
ts=torch.zeros((8,2048),dtype=torch.int64,device=dev)
_dbg(0, ts)
ops.implicit_gemm(x,w,pf,mf,sf,100000,masks,True,True)
torch.cuda.synchronize()
t=ts.cpu().numpy()
t0=t[3,0]
def rel(a): return [int(v-t0) for v in a if v>0]

prod=rel(t[0]); mma=rel(t[1]); epi=rel(t[2]); end=int(t[3,1]-t0)
n=len(prod)//2
print("kernel cycles (CTA0 role loop start -> final barrier):", end, " stages:", n, " tiles:", len(epi)//2)
f4=rel(t[4]); f5=[int(v-t0) for v in t[5][0::2] if v>0]
print("stage | prod: got_empty issued | mma: got_full fenced mma_issued committed loop_end")
for i in range(min(n,40)):
    print(f"{i:3d} | {prod[2*i]:7d} {prod[2*i+1]:7d} | {mma[2*i]:7d} {f4[2*i]:7d} {f4[2*i+1]:7d} {f5[i] if i < len(f5) else -1:7d} {mma[2*i+1]:7d}")
print("epilogue (start,end):", [(epi[2*i],epi[2*i+1]) for i in range(len(epi)//2)])
