"""perf triage: time fwd / dgrad / wgrad kernels of the bench workload under SPX_TC_DEBUG ablations"""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench_utils import surface_cloud
import spconv_b200.pytorch as spconv
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
x=torch.randn(100000,C,device=dev).half(); w=(torch.randn(K,3,3,3,C,device=dev)*0.05).half(); dout=torch.randn(100000,K,device=dev).half()
res=ops.get_indice_pairs_implicit_gemm(inds,1,shape,ConvAlgo.MaskImplicitGemm,[3]*3,[1]*3,[1]*3,[1]*3,[0]*3,True,False,is_train=True)
_,_,pf,pb,mf,mb,sf,sb,masks=res
flush=torch.empty(64<<20,device=dev)
def t(fn,n=20):
    for _ in range(3): fn()
    tot=0
    for _ in range(n):
        flush.zero_(); torch.cuda._sleep(600000); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot+=a.elapsed_time(b)
    return tot/n*1000
res_all={}
for ctas in ("1","2"):
    _dbg(0, None, int(ctas))
    out={}
    for dbg in [0,1,2,4,8,15]:
        _dbg(dbg, None, int(ctas))
        f=t(lambda: ops.implicit_gemm(x,w,pf,mf,sf,100000,masks,True,True))
        out[dbg]=round(f,1)
    _dbg(0, None, int(ctas))
    bw=t(lambda: ops.implicit_gemm_backward(x,w,dout,pf,pb,mf,mb,sf,sb,None,masks,128,True))
    res_all[ctas]={"fwd_us_by_debug":out,"bwd_us":round(bw,1)}
# host overhead of one op call (empty problem is not possible; time a tiny 128-row problem)
print(json.dumps(res_all))
