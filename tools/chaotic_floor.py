"""The chaotic floor of the free-running comparison, measured on the CPU with the oracle alone.

The 100-step IEMOCAP fixture (tests/golden/step_iemocap.pt) was produced by the UNMODIFIED reference in fp32.  Here the oracle
re-runs the same 100 steps free-running with ONE change that any correct implementation may differ by: the two encoder GEMMs
accumulate in fp64 (x.double() @ W.double().t(), rounded to fp32 afterwards) instead of the BLAS fp32 summation order.  Every
other operation, every input and every RNG draw is identical.  The drift of losses / weights / masks against the fixture is what
`argmax`, `score >= theta`, the `n >= 2` gates and Adam's m / (sqrt(v) + eps) on |g| ~ 1e-8 elements turn ulp-level differences
into -- the floor below which no free-running comparison of two implementations of this step can go.

Usage: python tools/chaotic_floor.py  > profiles/r02_chaotic_floor.log     (CPU only, ~1 minute)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F

import dad_b200  # noqa: F401
from helpers import load_golden, make_oracle, step_inputs
from oracle import dad_oracle as O

LOSS_KEYS = ("total_loss", "supervised_ce_loss", "consistency_loss", "ecda_loss")


def encoder_fp64_accumulate(x, pad, W, b):
    """O.encoder (I/model.py:18-41) with the GEMM accumulated in fp64; differentiable w.r.t. W, b like the original."""
    pre = (x.double() @ W.double().t() + b.double()).to(torch.float32)
    h = F.relu(pre)
    if pad is None:
        return h.mean(dim=1)
    v = 1 - pad.unsqueeze(-1).to(h.dtype)
    h = h * v
    n = (1 - pad.to(h.dtype)).sum(dim=1, keepdim=True)
    return h.sum(dim=1) / torch.clamp(n, min=1.0)


def run(variant):
    fix = load_golden("step_iemocap")
    meta = fix["meta"]
    orc, cfg, hp = make_oracle(meta["dataset"], meta["overrides"])
    real = O.encoder
    if variant == "fp64-accumulated GEMMs":
        O.encoder = encoder_fp64_accumulate
    worst, flips, k = 0.0, 0, 0
    try:
        for epoch in meta["epochs"]:
            orc.lr = O.cosine_lr(hp.LEARNING_RATE, epoch, hp.EPOCHS)
            for _ in range(meta["steps_per_epoch"]):
                c, n, rng = step_inputs(meta, k, cfg)
                got, ref = orc.step(c, n, epoch, rng), fix["steps"][k]
                worst = max(worst, max(abs(got[key] - ref[key]) / max(abs(ref[key]), 1e-3) for key in LOSS_KEYS))
                if epoch >= hp.WARMUP_EPOCHS:
                    flips += int((orc.last["mask"] != fix["mask"][k].bool()).sum())
                k += 1
            orc.end_epoch(epoch)
    finally:
        O.encoder = real
    sd = orc.state_dict()
    wd = max(float((sd[key] - ref).abs().max()) for key, ref in fix["final_state"].items())
    thr = float((orc.dacp.ema_thr - fix["dacp_ema_thr"]).abs().max())
    print(f"[{variant}] {k} free-running steps vs the reference fixture: worst loss rel = {worst:.3e}, max weight diff = {wd:.3e}, "
          f"mask flips = {flips}, threshold diff = {thr:.3e}")
    return worst, wd


if __name__ == "__main__":
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    print("chaotic floor of the 100-step free-running comparison (oracle on CPU; torch", torch.__version__, ")")
    l0, w0 = run("oracle unchanged, fp32 BLAS of THIS host (the fixture came from the reference on another host / BLAS kernel set)")
    l1, w1 = run("fp64-accumulated GEMMs")
    print(f"=> floor (summation order of the two encoder GEMMs alone): losses {max(l0, l1):.1e} relative, weights {max(w0, w1):.1e} absolute; "
          f"tests/test_gpu_step.py::test_iemocap_100_steps_free_running allows 6e-3 / 2.5e-3 (about 4x)")
