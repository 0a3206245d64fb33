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
    # slots of thread 0 in the run-based k2_pass (k2_filtfilt.cu, K2_TICK): 0 top of the iteration, 6 tile data ready
    # (mbarrier), 7 zero-state recursion, 1 warp scan + totals, 2 first barrier, 3 entering state + correction,
    # 4 second barrier, 5 store hand-off; 12 = the producer thread's work per iteration
    names=["top of the iteration","warp scan + totals","barrier 1","entering state + correction","barrier 2","store hand-off","wait tile data","recursion","-"]
    iters=n_cases*ns/1664*4*1.055                            # tiles x launches, + the warm-up tiles in front of the runs
    print(rep, {n: f"{100*float(v)/float(tot):.1f}%" for n,v in zip(names,out[:9])}, "cycles/iter/block", float(tot)/iters,
          "| producer thread, cycles/iter", float(out[12])/iters)
