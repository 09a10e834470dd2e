import numpy as np, sys
sys.path.insert(0,'/root/repo')
import kmer_count_b200 as K
from kmer_count_b200 import gen
world=2; n=300_000_000
ctxs=[K.KmerCounter(k=21,canonical=True,strategy=2) for _ in range(world)]
import torch
hists=[]
keep=[]
for r,kc in enumerate(ctxs):
    b=torch.empty(n,dtype=torch.uint8,device='cuda'); kc.gen_bases(5+r,0,n,b.data_ptr())
    off=torch.arange(0,n+1,400,dtype=torch.int64,device='cuda'); torch.cuda.synchronize()
    keep.append((b,off))
    kc.submit_device(b.data_ptr(),off.data_ptr(),n,off.numel()-1)
    h,low=kc.dist_hist(); hists.append(h)
ah=np.stack(hists)
for C in (1,6,8,12,16):
    need=ctxs[0].dist_plan(world,0,ah,C); print(C,need)
