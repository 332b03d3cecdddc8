"""Profiling driver (GPU): a few launches of the tcgen05 encoder forward / backward kernels at a bench-like size.
Usage: python tools/prof_enc.py [B] [T] [aug]   (aug: 0 clean, 1 weak/philox, 2 strong/philox)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dad_b200  # noqa: F401
from dad_b200 import _lib, ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
aug = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.manual_seed(0)
x = torch.randn(B, T, 768, device="cuda")
st = synth.init_state(42)
W, b = st["student_encoder.pre_net.weight"].cuda(), st["student_encoder.pre_net.bias"].cuda()
hi, lo = ops.split_f16(W)
keep = (torch.rand(768, device="cuda") > 0.1).float()
ts = torch.randint(0, T - int(T * 0.1) + 1, (B,), device="cuda", dtype=torch.int32)
a = None
if aug == 1:
    a = _lib.make_aug(1, sigma=0.01, seed=7, stream=1)
elif aug == 2:
    a = _lib.make_aug(2, sigma=0.05, feat_keep=keep, t_start=ts, mask_len=int(T * 0.1), seed=7, stream=2)
g = torch.randn(B, 256, device="cuda") * 1e-3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    e0.record()
    emb, bits, lens = ops.encoder_pool_fwd(x, None, W, b, aug=a, gemm_mode=_lib.GEMM_TC, W_hi=hi, W_lo=lo)
    e1.record(); torch.cuda.synchronize(); tf = e0.elapsed_time(e1)
    e0.record()
    dW, db = ops.encoder_pool_bwd(x, None, bits, g, lens, aug=a, gemm_mode=_lib.GEMM_TC)
    e1.record(); torch.cuda.synchronize(); tb = e0.elapsed_time(e1)
print(f"B={B} T={T} aug={aug}: fwd {tf:.3f} ms ({B*T*768*4/tf/1e6:.0f} GB/s) bwd {tb:.3f} ms ({B*T*768*4/tb/1e6:.0f} GB/s)")
