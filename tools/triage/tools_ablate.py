"""perf triage: time fwd / dgrad / wgrad kernels of the bench workload under SPX_TC_DEBUG ablations"""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench_utils import surface_cloud
import spconv_b200.pytorch as spconv
from spconv_b200.core import ConvAlgo
from spconv_b200.pytorch import ops
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
    os.environ["SPX_TC_CTAS"]=ctas
    out={}
    for dbg in [0,1,2,4,8,15]:
        os.environ["SPX_TC_DEBUG"]=str(dbg)
        f=t(lambda: ops.implicit_gemm(x,w,pf,mf,sf,100000,masks,True,True))
        out[dbg]=round(f,1)
    os.environ["SPX_TC_DEBUG"]="0"
    bw=t(lambda: ops.implicit_gemm_backward(x,w,dout,pf,pb,mf,mb,sf,sb,None,masks,128,True))
    res_all[ctas]={"fwd_us_by_debug":out,"bwd_us":round(bw,1)}
# host overhead of one op call (empty problem is not possible; time a tiny 128-row problem)
print(json.dumps(res_all))
