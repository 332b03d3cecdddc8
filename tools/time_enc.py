"""Timing table (GPU, CUDA events) of the tcgen05 encoder kernels alone at a bench-like size: every forward variant (one view /
fused two views, single CTA / CTA pair) and every backward variant (2 / 3 products, regenerated / cached B operand).
Usage: python tools/time_enc.py [B] [T] [reps]        prints one JSON object (ms per launch, kernel + k_pool_finalize / helpers)"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dad_b200  # noqa: F401
from dad_b200 import _lib, ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 5
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
xs = torch.zeros(2, B * T, 768, dtype=torch.float16, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(REPS):
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / REPS, 4)


out = {"B": B, "T": T, "reps": REPS, "alg_GB_per_view": B * T * 768 * 4 / 1e9}
state = {}
for pair in (0, 1):
    L.dad_tc_set_cta_pair(pair)
    tag = "pair" if pair else "single"
    out[f"fwd_clean_{tag}"] = timed(lambda: state.update(r=ops.encoder_pool_fwd(x, None, W, b, gemm_mode=_lib.GEMM_TC, W_hi=s_split[0], W_lo=s_split[1])))
    out[f"fwd_clean_cache_hi_{tag}"] = timed(lambda: ops.encoder_pool_fwd(x, None, W, b, gemm_mode=_lib.GEMM_TC, W_hi=s_split[0], W_lo=s_split[1], xs=(xs[0],)))
    out[f"fwd_weak_{tag}"] = timed(lambda: ops.encoder_pool_fwd(x, None, Wt, bt, aug=aw, gemm_mode=_lib.GEMM_TC, W_hi=t_split[0], W_lo=t_split[1], want_bits=False))
    out[f"fwd_strong_cache_hi_{tag}"] = timed(lambda: ops.encoder_pool_fwd(x, None, W, b, aug=as_, gemm_mode=_lib.GEMM_TC, W_hi=s_split[0], W_lo=s_split[1], xs=(xs[0],)))
    out[f"fwd_fused_cache_hi_{tag}"] = timed(lambda: ops.aug_encoder_fwd(x, None, Wt, bt, W, b, aw, as_, xs=(xs[0],), Wt_split=t_split, Ws_split=s_split))
    out[f"fwd_fused_nocache_{tag}"] = timed(lambda: ops.aug_encoder_fwd(x, None, Wt, bt, W, b, aw, as_, Wt_split=t_split, Ws_split=s_split))
    out[f"fwd_fused_cache_hilo_{tag}"] = timed(lambda: ops.aug_encoder_fwd(x, None, Wt, bt, W, b, aw, as_, xs=(xs[0], xs[1]), Wt_split=t_split, Ws_split=s_split))
L.dad_tc_set_cta_pair(2)  # auto: the training step's fused noisy forward runs on the transposed CTA-pair kernel (k_enc_fwd_tp)
out["fwd_fused_cache_hi_tp"] = timed(lambda: ops.aug_encoder_fwd(x, None, Wt, bt, W, b, aw, as_, xs=(xs[0],), Wt_split=t_split, Ws_split=s_split))
out["fwd_fused_nocache_tp"] = timed(lambda: ops.aug_encoder_fwd(x, None, Wt, bt, W, b, aw, as_, Wt_split=t_split, Ws_split=s_split))
out["fwd_fused_cache_hilo_tp"] = timed(lambda: ops.aug_encoder_fwd(x, None, Wt, bt, W, b, aw, as_, xs=(xs[0], xs[1]), Wt_split=t_split, Ws_split=s_split))
emb, bits, lens = state["r"]
for prods in (2, 3):
    L.dad_tc_set_bwd_products(prods)
    planes = (xs[0],) if prods == 2 else (xs[0], xs[1])
    ops.encoder_pool_fwd(x, None, W, b, gemm_mode=_lib.GEMM_TC, W_hi=s_split[0], W_lo=s_split[1], xs=planes)
    out[f"bwd_clean_regen_p{prods}"] = timed(lambda: ops.encoder_pool_bwd(x, None, bits, g, lens, gemm_mode=_lib.GEMM_TC))
    out[f"bwd_cached_p{prods}"] = timed(lambda: ops.encoder_pool_bwd(x, None, bits, g, lens, gemm_mode=_lib.GEMM_TC, xs=planes))
    out[f"bwd_strong_regen_p{prods}"] = timed(lambda: ops.encoder_pool_bwd(x, None, bits, g, lens, aug=as_, gemm_mode=_lib.GEMM_TC))
L.dad_tc_set_bwd_products(3)
L.dad_tc_set_cta_pair(2)
print(json.dumps(out))
