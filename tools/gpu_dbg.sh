#!/bin/bash
# Timing experiments on the transposed-pair fused forward (DAD_TC_DBG bits: 1 one product, 2 no W reloads, 4 no noise, 8 no A stores).
# Needs a library built with the experiment switches: DAD_NVCC_FLAGS=-DDAD_TC_EXPERIMENTS python <pkg>/csrc/build.py --force
OUT=gpurun_out; mkdir -p $OUT
for d in 0 1 2 4 8 3 6 12 15; do
  echo -n "dbg=$d "; DAD_TC_DBG=$d timeout 200 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, dad_b200
from dad_b200 import _lib, ops, synth
L=_lib.lib(); B,T=4096,200
torch.manual_seed(0)
x=torch.randn(B,T,768,device="cuda"); st=synth.init_state(42)
W,b=st["student_encoder.pre_net.weight"].cuda(),st["student_encoder.pre_net.bias"].cuda()
Wt,bt=(W*1.01).contiguous(),(b*0.99).contiguous()
ss,ts_=ops.split_f16(W),ops.split_f16(Wt)
keep=(torch.rand(768,device="cuda")>0.1).float()
ts=torch.randint(0,T-20+1,(B,),device="cuda",dtype=torch.int32)
aw=_lib.make_aug(1,sigma=0.01,seed=7,stream=_lib.STREAM_WEAK)
as_=_lib.make_aug(2,sigma=0.05,feat_keep=keep,t_start=ts,mask_len=20,seed=7,stream=_lib.STREAM_STRONG)
xs=torch.zeros(2,B*T,768,dtype=torch.float16,device="cuda")
L.dad_tc_set_cta_pair(2)
f=lambda: ops.aug_encoder_fwd(x,None,Wt,bt,W,b,aw,as_,xs=(xs[0],xs[1]),Wt_split=ts_,Ws_split=ss)
for _ in range(3): f()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
tot=0
for _ in range(10):
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); tot+=e0.elapsed_time(e1)
print(round(tot/10,4))
PY
done
