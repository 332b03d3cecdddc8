"""Profiling driver (GPU) for the HBM-bound kernels the north star names (augmentation, EMA / Adam, split-K fold, pooling finalize,
device-side batch assembly): one launch each at the headline size, for `ncu --set full` (achieved DRAM GB/s).
Usage: python tools/prof_small.py [B] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dad_b200  # noqa: F401
from dad_b200 import _lib, config as cfg, ops, synth
from dad_b200.engine import DADStepEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda")
x = torch.randn(B, T, 768, device=dev)
keep = (torch.rand(768, device=dev) > 0.1).float()
ts = torch.randint(0, T - int(T * 0.1) + 1, (B,), device=dev, dtype=torch.int32)
for _ in range(2):
    w = ops.augment(x, 0.01, seed=5, stream=_lib.STREAM_WEAK)                                                     # k_augment<true> (weak)
    s = ops.augment(x, 0.05, seed=5, stream=_lib.STREAM_STRONG, feat_keep=keep, t_start=ts, mask_len=int(T * 0.1), strong=True)
    n = ops.augment(x, 0.05, noise=w, feat_keep=keep, t_start=ts, mask_len=int(T * 0.1), strong=True)               # supplied noise
del w, s, n
feats = x.reshape(B * T, 768)
offs = (torch.arange(B, device=dev) * T).to(torch.int64)
sizes = torch.randint(T // 2, T + 1, (B,), device=dev, dtype=torch.int32)
for _ in range(2):
    ops.gather_pad(feats, offs, sizes, T=T)                                                                         # k_gather_pad
eng = DADStepEngine(cfg, B, T, gemm_mode=_lib.GEMM_TC, use_graph=False, noise_mode="philox")
eng.load_state_dict(synth.init_state(42))
eng.set_inputs({"net_input": {"feats": x, "padding_mask": None}, "labels": torch.randint(0, 4, (B,), device=dev)},
               {"net_input": {"feats": x, "padding_mask": None}})
eng.set_lr(1e-3)
for _ in range(2):
    eng.step(0)
eng.init_teacher()
for _ in range(2):
    eng.step(60)                                                                                                   # k_pool_finalize, k_grad_reduce, k_grad_norm, k_adam_ema
torch.cuda.synchronize()
print("ok")
