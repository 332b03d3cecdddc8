"""Dev check (GPU): tcgen05 encoder kernels vs the exact-fp32 CUDA-core kernels on the same inputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dad_b200
from dad_b200 import _lib, ops, synth

def run(B, T, aug_mode, do_bwd):
    torch.manual_seed(0)
    dev = "cuda"
    c, n = synth.make_batches(B, T, 5, variable_len=True)
    rng = synth.make_rng_inputs(B, T, 6)
    st = synth.init_state(42)
    W, b = st["student_encoder.pre_net.weight"].cuda(), st["student_encoder.pre_net.bias"].cuda()
    x, pad = n["net_input"]["feats"].cuda(), n["net_input"]["padding_mask"].cuda()
    keep = (rng["feat_keep_u"] > 0.1).float().cuda(); ts = rng["t_start"].to(torch.int32).cuda()
    nw, ns = rng["n_weak"].cuda(), rng["n_strong"].cuda()
    L = int(T * 0.1)
    def mk():
        if aug_mode == 0: return None
        if aug_mode == 1: return _lib.make_aug(1, noise=nw, sigma=0.01)
        if aug_mode == 2: return _lib.make_aug(2, noise=ns, sigma=0.05, feat_keep=keep, t_start=ts, mask_len=L)
        return _lib.make_aug(2, sigma=0.05, feat_keep=keep, t_start=ts, mask_len=L, seed=11, stream=2)
    e0, bits0, lens0 = ops.encoder_pool_fwd(x, pad, W, b, aug=mk(), gemm_mode=_lib.GEMM_FP32)
    torch.cuda.synchronize()
    e1, bits1, lens1 = ops.encoder_pool_fwd(x, pad, W, b, aug=mk(), gemm_mode=_lib.GEMM_TC)
    torch.cuda.synchronize()
    d = (e0 - e1).abs().max().item(); s = e0.abs().max().item()
    diffbits = (bits0 != bits1).sum().item()
    print(f"B={B} T={T} aug={aug_mode}: emb max abs diff {d:.3e} (scale {s:.3e}) rel {d/s:.2e}; lens equal {torch.equal(lens0, lens1)}; differing bit-words {diffbits}/{bits0.numel()}")
    if do_bwd:
        g = torch.randn(B, 256, device=dev)
        dW0, db0 = ops.encoder_pool_bwd(x, pad, bits0, g, lens0, aug=mk(), gemm_mode=_lib.GEMM_FP32)
        dW1, db1 = ops.encoder_pool_bwd(x, pad, bits0, g, lens0, aug=mk(), gemm_mode=_lib.GEMM_TC)
        torch.cuda.synchronize()
        print(f"   dW rel-to-max diff {((dW0-dW1).abs().max()/dW0.abs().max()).item():.2e}  db {((db0-db1).abs().max()/db0.abs().max()).item():.2e}")

if __name__ == "__main__":
    do_bwd = "--bwd" in sys.argv
    for (B, T, a) in [(2, 64, 0), (4, 200, 0), (3, 97, 1), (5, 200, 2), (64, 200, 3), (7, 33, 2), (300, 211, 0)]:
        run(B, T, a, do_bwd)
    # timing
    B, T = 512, 200
    x = torch.randn(B, T, 768, device="cuda"); st = synth.init_state(42)
    W, b = st["student_encoder.pre_net.weight"].cuda(), st["student_encoder.pre_net.bias"].cuda()
    hi, lo = ops.split_f16(W)
    for mode in (_lib.GEMM_FP32, _lib.GEMM_TC):
        for _ in range(3): ops.encoder_pool_fwd(x, None, W, b, gemm_mode=mode, W_hi=hi, W_lo=lo)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.encoder_pool_fwd(x, None, W, b, gemm_mode=mode, W_hi=hi, W_lo=lo)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"fwd mode {mode}: {ms:.3f} ms  {B*T*768*4/ms/1e6:.0f} GB/s  {2*B*T*768*256/ms/1e9:.1f} TFLOP/s (algorithmic)")
