"""perf triage: SubM rulebook (memset + insert + probe) with optional outputs switched off"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench_utils import surface_cloud
from spconv_b200 import _cabi
from spconv_b200.pytorch import ops
dev = torch.device("cuda:0")
shape=[41,1600,1408]
rng=np.random.default_rng(50051)
N=100000
lib=ops._lib()
for order in ("shuffled","sorted"):
    inds_np=surface_cloud(rng,shape,N)
    if order=="sorted":
        key=((inds_np[:,1].astype(np.int64)*shape[1]+inds_np[:,2])*shape[2]+inds_np[:,3])
        inds_np=inds_np[np.argsort(key)]
    inds=torch.from_numpy(inds_np).to(dev)
    out_shape=list(shape)
    geo=ops._geometry(inds,1,shape,out_shape,[3]*3,[1]*3,[1]*3,[1]*3,False)
    pf=torch.empty((27,N),dtype=torch.int32,device=dev); pb=torch.empty_like(pf)
    mask=torch.empty((N,),dtype=torch.int32,device=dev); rows=torch.empty((N,32),dtype=torch.int32,device=dev)
    ws=torch.empty(lib.spx_rulebook_workspace_size(ctypes.byref(geo),N,0,1),dtype=torch.uint8,device=dev)
    flush=torch.empty(64<<20,device=dev)
    def run(pb_,mask_,rows_):
        def f():
            _cabi.check(lib.spx_subm_rulebook(ctypes.byref(geo),inds.data_ptr(),N,pf.data_ptr(),pb_,mask_,rows_,ws.data_ptr(),ws.numel(),ops._stream()),"x")
        for _ in range(3): f()
        tot=0
        for _ in range(10):
            flush.zero_(); torch.cuda._sleep(400000)
            a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize(); tot+=a.elapsed_time(b)
        return tot/10*1000
    print(order,"all outputs            :", round(run(pb.data_ptr(),mask.data_ptr(),rows.data_ptr()),1),"us")
    print(order,"no row table           :", round(run(pb.data_ptr(),mask.data_ptr(),None),1),"us")
    print(order,"no row table, no bwd   :", round(run(None,mask.data_ptr(),None),1),"us")
    print(order,"pair_fwd only          :", round(run(None,None,None),1),"us")
