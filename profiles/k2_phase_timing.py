import os, sys, ctypes, numpy as np
os.environ["AEW_K2_TIMING"]="1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched, _cabi
dev=torch.device('cuda')
n_cases=int(os.environ.get("NC","200")); ns=1800000
flat=bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off=np.arange(n_cases+1,dtype=np.int64)*ns
batch=batched.RecordBatch(flat, rec_off)
pre=batched.Preprocessor()
beg_h,end_h,d_beg,d_end=pre._ranges(batch, trim=False)
mean,sd,cnt=pre.zscore_stats(flat,beg_h,end_h,d_beg,d_end)
y=torch.empty_like(flat); tmp=torch.empty_like(flat)
lib=_cabi.load()
out=np.zeros(16,np.uint64)
for rep in range(3):
    pre.filter_records(flat,beg_h,end_h,d_beg,d_end,mean,sd,cnt,out=y,tmp=tmp)
    lib.aew_filt_timing_dump(ctypes.c_void_p(out.ctypes.data))
    tot=out[:9].sum()
    names=["early look-back loads","warp scan","barrier after A","publish+lookback+chains","barrier 2","B half+store","wait tile data","phase 1 recursion","claim+describe+bulk issue"]
    print(rep, {n: f"{100*float(v)/float(tot):.1f}%" for n,v in zip(names,out[:9])}, "cycles/iter/block", float(tot)/(n_cases*ns/3328*4), "look-back steps that had to wait:", int(out[9]), "of", int(out[10]), "| cycles/iter: look-back thread", float(out[11])/(n_cases*ns/3328*4), "producer thread", float(out[12])/(n_cases*ns/3328*4),
          "| thread 0 inside the serial section: scan", float(out[13])/(n_cases*ns/3328*4), "store + publish", float(out[14])/(n_cases*ns/3328*4))
