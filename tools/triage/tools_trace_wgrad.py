"""perf triage: clock64 timeline of CTA (0,0) of the weight-gradient kernel (SPX_TC_TRACE)"""
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
x=torch.randn(100000,C,device=dev).half(); w=(torch.randn(K,3,3,3,C,device=dev)*0.05).half(); dout=torch.randn(100000,K,device=dev).half()
res=ops.get_indice_pairs_implicit_gemm(inds,1,shape,ConvAlgo.MaskImplicitGemm,[3]*3,[1]*3,[1]*3,[1]*3,[0]*3,True,False,is_train=True)
_,_,pf,pb,mf,mb,sf,sb,masks=res
bw=lambda: ops.implicit_gemm_backward(x,w,dout,pf,pb,mf,mb,sf,sb,None,masks,128,True)
for _ in range(3): bw()
ts=torch.zeros((8,2048),dtype=torch.int64,device=dev)
_dbg(int(os.environ.get("WG_DEBUG","0")), ts)
bw(); torch.cuda.synchronize()
t=ts.cpu().numpy(); t0=t[3,0]
rel=lambda a:[int(v-t0) for v in a if v>0]
prod=rel(t[0]); mma=rel(t[1]); pb_=rel(t[6]); mb_=rel(t[7])
print("loop end / after final sync:", int(t[3,1]-t0), int(t[3,2]-t0), " stages:", len(prod)//2, " tiles:", len(pb_)//2)
print("tile | prod: got_empty_b issued_b | mma got_full_b")
for i in range(min(len(pb_)//2,14)): print(f"{i:3d} | {pb_[2*i]:7d} {pb_[2*i+1]:7d} | {mb_[i] if i<len(mb_) else -1:7d}")
t5=rel(t[5])
print("tile top | after fetch | after idx_full | after active_groups")
for i in range(min(len(t5)//4,14)): print(i, t5[4*i:4*i+4])
print("stage | prod: got_empty_a issued | mma: got_full_a issued+committed")
for i in range(min(len(prod)//2,36)): print(f"{i:3d} | {prod[2*i]:7d} {prod[2*i+1]:7d} | {mma[2*i]:7d} {mma[2*i+1]:7d}")

sp=t[2][:512*4].reshape(-1,4); sp=sp[sp[:,0]>0]
if len(sp):
    s0=sp[:,0].min(); r=(sp-s0)/1000.0
    half=len(r)//2
    for name,rr in (("pass 0",r[:half]),("pass 1",r[half:])):
        d=rr[:,2]-rr[:,1]
        print(f"{name}: CTAs {len(rr)}  main loop (prologue done -> accumulators done) min {d.min():.1f} median {np.median(d):.1f} max {d.max():.1f} us; exit median {np.median(rr[:,3]):.1f} max {rr[:,3].max():.1f} us")
