"""Profiling driver (GPU): ONE launch of each tcgen05 encoder forward / backward variant at a bench-like size, for ncu.
Usage: python tools/prof_enc.py [B] [T] [which]     which = comma list of: clean, fused, weak, strong, bwd_clean, bwd_cached
                                                     suffix ':single' runs the forward variants without CTA pairs, ':tp' with the default policy (fused -> k_enc_fwd_tp)
e.g.   ncu --set full --import-source on -k regex:k_enc_ -o gpurun_out/enc python tools/prof_enc.py 4096 200 clean,fused"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dad_b200  # noqa: F401
from dad_b200 import _lib, ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
which = (sys.argv[3] if len(sys.argv) > 3 else "clean,fused,bwd_clean,bwd_cached").split(",")
L = _lib.lib()
torch.manual_seed(0)
x = torch.randn(B, T, 768, device="cuda")
st = synth.init_state(42)
W, b = st["student_encoder.pre_net.weight"].cuda(), st["student_encoder.pre_net.bias"].cuda()
Wt, bt = (W * 1.01).contiguous(), (b * 0.99).contiguous()
s_split, t_split = ops.split_f16(W), ops.split_f16(Wt)
keep = (torch.rand(768, device="cuda") > 0.1).float()
ts = torch.randint(0, T - int(T * 0.1) + 1, (B,), device="cuda", dtype=torch.int32)
aw = _lib.make_aug(1, sigma=0.01, seed=7, stream=_lib.STREAM_WEAK)
as_ = _lib.make_aug(2, sigma=0.05, feat_keep=keep, t_start=ts, mask_len=int(T * 0.1), seed=7, stream=_lib.STREAM_STRONG)
g = torch.randn(B, 256, device="cuda") * 1e-3
xs = torch.zeros(B * T, 768, dtype=torch.float16, device="cuda")
emb, bits, lens = ops.encoder_pool_fwd(x, None, W, b, gemm_mode=_lib.GEMM_TC, W_hi=s_split[0], W_lo=s_split[1], xs=(xs,))
torch.cuda.synchronize()
for item in which:
    name, _, flavour = item.partition(":")
    L.dad_tc_set_cta_pair({"single": 0, "tp": 2, "auto": 2}.get(flavour, 1))  # 'fused:tp' = the transposed CTA-pair kernel
    if name == "clean":
        ops.encoder_pool_fwd(x, None, W, b, gemm_mode=_lib.GEMM_TC, W_hi=s_split[0], W_lo=s_split[1])
    elif name == "weak":
        ops.encoder_pool_fwd(x, None, Wt, bt, aug=aw, gemm_mode=_lib.GEMM_TC, W_hi=t_split[0], W_lo=t_split[1], want_bits=False)
    elif name == "strong":
        ops.encoder_pool_fwd(x, None, W, b, aug=as_, gemm_mode=_lib.GEMM_TC, W_hi=s_split[0], W_lo=s_split[1], xs=(xs,))
    elif name == "fused":
        ops.aug_encoder_fwd(x, None, Wt, bt, W, b, aw, as_, xs=(xs,), Wt_split=t_split, Ws_split=s_split)
    elif name == "bwd_clean":
        ops.encoder_pool_bwd(x, None, bits, g, lens, gemm_mode=_lib.GEMM_TC)
    elif name == "bwd_cached":
        ops.encoder_pool_bwd(x, None, bits, g, lens, gemm_mode=_lib.GEMM_TC, xs=(xs,))
    else:
        raise SystemExit(f"unknown variant {name}")
    torch.cuda.synchronize()
print("ok", which)
