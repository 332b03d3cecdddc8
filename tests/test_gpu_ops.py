"""GPU (-m gpu): op-level parity of the CUDA kernels, called through the C ABI, against the oracle and the
reference-generated golden vectors.  Augmentation must be BIT-EXACT; floating-point ops within the stated tolerances."""
import os
import sys

import pytest
import torch

from helpers import load_golden, pretrained_state
from dad_b200 import synth

pytestmark = pytest.mark.gpu
MODES = ["fp32", "tc"]


def _mode(name):
    from dad_b200 import _lib
    return {"fp32": _lib.GEMM_FP32, "tc": _lib.GEMM_TC}[name]


def _inputs(B, T, seed, variable=True):
    c, n = synth.make_batches(B, T, seed, variable_len=variable)
    rng = synth.make_rng_inputs(B, T, seed + 1)
    return c, n, rng


def test_library_loaded_and_device_ok():
    from dad_b200 import _lib
    _lib.require_device()
    assert _lib.lib().dad_abi_version() == 3


@pytest.fixture
def tc_switch():
    """Set the process-wide tcgen05 switches (CTA pair on/off, products of the dW GEMM) for one test and restore them."""
    from dad_b200 import _lib
    L = _lib.lib()
    old = (L.dad_tc_get_cta_pair(), L.dad_tc_get_bwd_products())

    def set_(cta_pair=None, products=None):  # cta_pair: 0 never, 1 always, 2 auto (bool accepted)
        if cta_pair is not None:
            L.dad_tc_set_cta_pair(int(cta_pair))
        if products is not None:
            L.dad_tc_set_bwd_products(int(products))
    yield set_
    L.dad_tc_set_cta_pair(old[0])
    L.dad_tc_set_bwd_products(old[1])


@pytest.mark.parametrize("B,T", [(6, 20), (3, 7), (2, 200), (1, 1)])
def test_augment_bit_exact_vs_oracle(B, T):
    from dad_b200 import ops
    from oracle import dad_oracle as O
    c, n, rng = _inputs(B, T, 777)
    x = n["net_input"]["feats"]
    dev = "cuda"
    weak = ops.augment(x.to(dev), 0.01, noise=rng["n_weak"].to(dev))
    assert torch.equal(weak.cpu(), O.weak_augment(x, rng["n_weak"], 0.01))
    L = int(T * 0.1)
    keep = (rng["feat_keep_u"] > 0.1).float()
    strong = ops.augment(x.to(dev), 0.05, noise=rng["n_strong"].to(dev), feat_keep=keep.to(dev),
                         t_start=rng["t_start"].to(torch.int32).to(dev) if L > 0 else None, mask_len=L, strong=True)
    ref = O.strong_augment(x, rng["n_strong"], 0.05, rng["feat_keep_u"], 0.1, rng["t_start"], 0.1)
    # bit-exact including the sign of zeros (x*0 -> +-0, span := +0)
    assert torch.equal(strong.cpu().view(torch.int32), ref.view(torch.int32))


def test_augment_bit_exact_vs_reference_golden():
    from dad_b200 import ops
    g = load_golden("ops_small")
    c, n = synth.make_batches(g["B"], g["T"], g["seed_batches"], variable_len=True)
    rng = synth.make_rng_inputs(g["B"], g["T"], g["seed_rng"])
    x = n["net_input"]["feats"].cuda()
    weak = ops.augment(x, 0.01, noise=rng["n_weak"].cuda())
    assert torch.equal(weak.cpu().view(torch.int32), g["weak"].view(torch.int32))
    keep = (rng["feat_keep_u"] > 0.1).float().cuda()
    strong = ops.augment(x, 0.05, noise=rng["n_strong"].cuda(), feat_keep=keep, t_start=rng["t_start"].to(torch.int32).cuda(),
                         mask_len=int(g["T"] * 0.1), strong=True)
    assert torch.equal(strong.cpu().view(torch.int32), g["strong"].view(torch.int32))
    s2 = ops.augment(x[0], 0.05, noise=rng["n_strong"][0].cuda(), feat_keep=keep, t_start=rng["t_start"][:1].to(torch.int32).cuda(),
                     mask_len=int(g["T"] * 0.1), strong=True)
    assert torch.equal(s2.cpu().view(torch.int32), g["strong2d"].view(torch.int32))


def test_augment_philox_statistics_and_determinism():
    from dad_b200 import ops
    x = torch.zeros(8, 64, 768, device="cuda")
    a = ops.augment(x, 1.0, seed=5, stream=2)
    b = ops.augment(x, 1.0, seed=5, stream=2)
    c = ops.augment(x, 1.0, seed=5, stream=3)   # the other half of the same Philox blocks (dad_common.cuh): independent normals
    d = ops.augment(x, 1.0, seed=5, stream=4)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    assert abs(float((a * c).mean())) < 5e-3 and abs(float((a * d).mean())) < 5e-3
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1.0) < 5e-3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05  # normal kurtosis


@pytest.mark.parametrize("mode", MODES)
def test_encoder_forward_vs_reference_golden(mode):
    from dad_b200 import ops
    g = load_golden("ops_small")
    c, n = synth.make_batches(g["B"], g["T"], g["seed_batches"], variable_len=True)
    st = pretrained_state()
    W, b = st["student_encoder.pre_net.weight"].cuda(), st["student_encoder.pre_net.bias"].cuda()
    if mode == "tc":
        pytest.skip("T=20 < 32: tcgen05 path not applicable to this fixture")
    x, pad = c["net_input"]["feats"].cuda(), c["net_input"]["padding_mask"].cuda()
    emb, bits, lens = ops.encoder_pool_fwd(x, pad, W, b, gemm_mode=_mode(mode))
    assert torch.allclose(emb.cpu(), g["emb_pad"], rtol=2e-5, atol=2e-6), (emb.cpu() - g["emb_pad"]).abs().max()
    emb2, _, lens2 = ops.encoder_pool_fwd(x, None, W, b, gemm_mode=_mode(mode))
    assert torch.allclose(emb2.cpu(), g["emb_nopad"], rtol=2e-5, atol=2e-6)
    assert torch.equal(lens.cpu(), (~c["net_input"]["padding_mask"]).sum(1).float())
    assert torch.equal(lens2.cpu(), torch.full((g["B"],), float(g["T"])))


# tcgen05 variants: default (CTA pairs where they pay, 3-product dW GEMM), pairs forced on / off, the opt-in 2-product dW GEMM
TC_VARIANTS = ["fp32", "tc", "tc-single-cta", "tc-pair", "tc-2prod"]


@pytest.mark.parametrize("mode", TC_VARIANTS)
@pytest.mark.parametrize("B,T,aug", [(4, 64, 0), (5, 200, 2), (3, 97, 1), (2, 33, 2), (16, 40, 2), (1, 1, 0), (2, 5, 2), (41, 100, 2)])
def test_encoder_forward_backward_vs_oracle(mode, B, T, aug, tc_switch):
    """emb, ReLU bits, dW, db against the oracle's autograd on the same (augmented) inputs."""
    from dad_b200 import _lib, ops
    from oracle import dad_oracle as O
    if mode != "fp32" and T < 32:
        pytest.skip("tcgen05 path needs T >= 32")
    tc_switch(cta_pair={"tc-single-cta": 0, "tc-pair": 1}.get(mode, 2), products=2 if mode == "tc-2prod" else 3)
    mode = "fp32" if mode == "fp32" else "tc"
    c, n, rng = _inputs(B, T, 4242 + B * 1000 + T)
    st = pretrained_state()
    W0, b0 = st["student_encoder.pre_net.weight"], st["student_encoder.pre_net.bias"]
    x, pad = n["net_input"]["feats"], n["net_input"]["padding_mask"]
    L = int(T * 0.1)
    keep = (rng["feat_keep_u"] > 0.1).float()
    if aug == 0:
        xa, a = x, None
    elif aug == 1:
        xa = O.weak_augment(x, rng["n_weak"], 0.01)
        noise_d = rng["n_weak"].cuda()  # keep alive: dad_aug only holds the raw pointer
        a = _lib.make_aug(1, noise=noise_d, sigma=0.01)
    else:
        xa = O.strong_augment(x, rng["n_strong"], 0.05, rng["feat_keep_u"], 0.1, rng["t_start"], 0.1)
        keep_d, ts_d, noise_d = keep.cuda(), rng["t_start"].to(torch.int32).cuda(), rng["n_strong"].cuda()
        a = _lib.make_aug(2, noise=noise_d, sigma=0.05, feat_keep=keep_d, t_start=ts_d if L > 0 else None, mask_len=L)
    W, b = W0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    emb_ref = O.encoder(xa, pad, W, b)
    gen = torch.Generator().manual_seed(9)
    g_emb = torch.randn(B, 256, generator=gen)
    (emb_ref * g_emb).sum().backward()
    emb, bits, lens = ops.encoder_pool_fwd(x.cuda(), pad.cuda(), W0.cuda(), b0.cuda(), aug=a, gemm_mode=_mode(mode))
    tol = dict(rtol=2e-5, atol=2e-6) if mode == "fp32" else dict(rtol=1e-4, atol=2e-5)
    assert torch.allclose(emb.cpu(), emb_ref.detach(), **tol), (emb.cpu() - emb_ref.detach()).abs().max()
    # ReLU sign bits against fp64 pre-activations (ignore |pre| < 1e-4: rounding may legitimately flip those)
    pre = (xa.double() @ W0.double().t() + b0.double()).reshape(B * T, 256)
    word = bits.cpu().view(B * T, 8).to(torch.int64) & 0xFFFFFFFF
    got = ((word.unsqueeze(-1) >> torch.arange(32)) & 1).reshape(B * T, 256).bool()
    valid_f = (~pad).reshape(B * T, 1)
    sure = (pre.abs() > 1e-4) & valid_f
    assert torch.equal(got[sure], (pre > 0)[sure])
    assert not bool(got[~valid_f.expand(-1, 256)].any())  # padded frames carry no ReLU bits (they never reach the backward)
    dW, db = ops.encoder_pool_bwd(x.cuda(), pad.cuda(), bits, g_emb.cuda(), lens, aug=a, gemm_mode=_mode(mode))
    sW, sb = float(W.grad.abs().max()), float(b.grad.abs().max())
    tolW = 2e-5 if mode == "fp32" else 1e-4
    assert float((dW.cpu() - W.grad).abs().max()) <= tolW * sW + 1e-7, (float((dW.cpu() - W.grad).abs().max()), sW)
    assert float((db.cpu() - b.grad).abs().max()) <= tolW * sb + 1e-7


def test_fp32_encoder_workspace_reuse_across_batch_sizes():
    """The op workspace is cached and reused across calls, and the fp32 kernel's arrival counters occupy its first 2*B words:
    a larger batch must not inherit the pooled partials a smaller batch left in those words (regression: lens / emb garbage at
    B=4096 after a B=512 call; dad_encoder_fwd now zeroes the counters itself)."""
    from dad_b200 import _lib, ops
    from oracle import dad_oracle as O
    st = pretrained_state()
    W0, b0 = st["student_encoder.pre_net.weight"], st["student_encoder.pre_net.bias"]
    for B in (8, 100, 3, 300, 100):
        c, _, _ = _inputs(B, 40, 99 + B)
        x, pad = c["net_input"]["feats"], c["net_input"]["padding_mask"]
        emb, _, lens = ops.encoder_pool_fwd(x.cuda(), pad.cuda(), W0.cuda(), b0.cuda(), gemm_mode=_lib.GEMM_FP32)
        assert torch.equal(lens.cpu(), (~pad).sum(1).float()), B
        assert torch.allclose(emb.cpu(), O.encoder(x, pad, W0, b0), rtol=2e-5, atol=2e-6), B


@pytest.mark.parametrize("products", [2, 3])
@pytest.mark.parametrize("B,T,aug", [(4, 64, 0), (5, 200, 2), (3, 97, 1), (2, 33, 2), (37, 41, 2)])
def test_operand_cache_backward_is_bit_identical(B, T, aug, products, tc_switch):
    """tcgen05 mode: the backward fed by TMA from the forward's fp16 operand cache (hi only for the 2-product dW GEMM, hi and lo
    for the 3-product one) == the backward that re-reads, re-augments and re-splits x (same operand values -> same dW partials,
    bit for bit); emb / bits are unaffected."""
    from dad_b200 import _lib, ops
    tc_switch(products=products)
    c, n, rng = _inputs(B, T, 777 + B * 1000 + T)
    st = pretrained_state()
    W, b = st["student_encoder.pre_net.weight"].cuda(), st["student_encoder.pre_net.bias"].cuda()
    x, pad = n["net_input"]["feats"].cuda(), n["net_input"]["padding_mask"].cuda()
    L = int(T * 0.1)
    a = None
    if aug == 1:
        noise_d = rng["n_weak"].cuda()
        a = _lib.make_aug(1, noise=noise_d, sigma=0.01)
    elif aug == 2:
        keep_d, ts_d, noise_d = (rng["feat_keep_u"] > 0.1).float().cuda(), rng["t_start"].to(torch.int32).cuda(), rng["n_strong"].cuda()
        a = _lib.make_aug(2, noise=noise_d, sigma=0.05, feat_keep=keep_d, t_start=ts_d if L > 0 else None, mask_len=L)
    g_emb = torch.randn(B, 256, generator=torch.Generator().manual_seed(11)).cuda()
    emb0, bits0, lens0 = ops.encoder_pool_fwd(x, pad, W, b, aug=a, gemm_mode=_lib.GEMM_TC)
    dW0, db0 = ops.encoder_pool_bwd(x, pad, bits0, g_emb, lens0, aug=a, gemm_mode=_lib.GEMM_TC)
    xs = torch.full((products - 1, B * T, 768), float("nan"), dtype=torch.float16, device="cuda")
    planes = tuple(xs[i] for i in range(products - 1))
    emb1, bits1, lens1 = ops.encoder_pool_fwd(x, pad, W, b, aug=a, gemm_mode=_lib.GEMM_TC, xs=planes)
    assert torch.equal(emb0, emb1) and torch.equal(bits0, bits1)
    assert not torch.isnan(xs.float()).any()  # every frame row of the cache was written
    dW1, db1 = ops.encoder_pool_bwd(x, pad, bits1, g_emb, lens1, aug=None, gemm_mode=_lib.GEMM_TC, xs=planes)
    assert torch.equal(dW0, dW1) and torch.equal(db0, db1)
    # the cache holds the augmented view itself: hi (+ lo) == 8 * aug(x) to 2^-12 (2^-21)
    if aug == 0:
        rec = sum(pl.float() for pl in planes) / 8.0
        tol = 2e-6 if products == 3 else 5e-4  # fp16 half-ulp of the top binade
        assert float((rec.view(B, T, 768) - x).abs().max()) <= tol * float(x.abs().max())


def _noisy_views(B, T, seed, philox):
    from dad_b200 import _lib
    c, n, rng = _inputs(B, T, seed)
    x, pad = n["net_input"]["feats"].cuda(), n["net_input"]["padding_mask"].cuda()
    L = int(T * 0.1)
    keep_d, ts_d = (rng["feat_keep_u"] > 0.1).float().cuda(), rng["t_start"].to(torch.int32).cuda()
    hold = [keep_d, ts_d]
    if philox:
        aw = _lib.make_aug(1, sigma=0.01, seed=77, stream=_lib.STREAM_WEAK)
        as_ = _lib.make_aug(2, sigma=0.05, feat_keep=keep_d, t_start=ts_d if L > 0 else None, mask_len=L, seed=77, stream=_lib.STREAM_STRONG)
    else:
        nw, ns = rng["n_weak"].cuda(), rng["n_strong"].cuda()
        hold += [nw, ns]
        aw = _lib.make_aug(1, noise=nw, sigma=0.01)
        as_ = _lib.make_aug(2, noise=ns, sigma=0.05, feat_keep=keep_d, t_start=ts_d if L > 0 else None, mask_len=L)
    return x, pad, aw, as_, hold, (n, rng)


@pytest.mark.parametrize("cta_pair", [2, 1, 0], ids=["transposed-pair", "pair", "single"])
@pytest.mark.parametrize("philox", [False, True], ids=["supplied", "philox"])
@pytest.mark.parametrize("B,T", [(4, 64), (5, 200), (3, 97), (2, 33), (37, 41), (64, 200), (9, 128)])
def test_fused_noisy_forward_equals_the_two_single_view_forwards(B, T, philox, cta_pair, tc_switch):
    """dad_aug_encoder_fwd (I/train.py:405-409,439 as one kernel, x_noisy read once) == dad_encoder_fwd of the weak view on the
    teacher weights + dad_encoder_fwd of the strong view on the student weights: embeddings, ReLU bits and the operand cache
    bit for bit (same operands, same accumulation order) -- with supplied noise and with the in-kernel Philox field (whose weak /
    strong streams share their Philox blocks).  cta_pair = 2 (the default policy) runs the fused launch on the transposed CTA-pair
    kernel k_enc_fwd_tp (weights as the A operand, frames as B): same products in the same order, so the ReLU bits and the operand
    cache stay bit-identical; its pooled sums add the frames of a 32-frame group serially instead of by a butterfly, so the
    embeddings agree to fp32 rounding of that sum."""
    from dad_b200 import _lib, ops
    tc_switch(cta_pair=cta_pair)
    st = pretrained_state()
    Ws, bs = st["student_encoder.pre_net.weight"].cuda(), st["student_encoder.pre_net.bias"].cuda()
    g = torch.Generator().manual_seed(5)
    Wt = (Ws.cpu() + 0.01 * torch.randn(256, 768, generator=g)).cuda()   # a teacher that differs from the student
    bt = (bs.cpu() + 0.01 * torch.randn(256, generator=g)).cuda()
    x, pad, aw, as_, hold, _ = _noisy_views(B, T, 31 + B * 1000 + T, philox)
    emb_t0, _, lens0 = ops.encoder_pool_fwd(x, pad, Wt, bt, aug=aw, gemm_mode=_lib.GEMM_TC, want_bits=False)
    xs0 = torch.full((B * T, 768), float("nan"), dtype=torch.float16, device="cuda")
    emb_s0, bits0, _ = ops.encoder_pool_fwd(x, pad, Ws, bs, aug=as_, gemm_mode=_lib.GEMM_TC, xs=(xs0,))
    xs1 = torch.full((B * T, 768), float("nan"), dtype=torch.float16, device="cuda")
    emb_t1, emb_s1, bits1, lens1 = ops.aug_encoder_fwd(x, pad, Wt, bt, Ws, bs, aw, as_, xs=(xs1,))
    if cta_pair == 2:
        assert torch.allclose(emb_t0, emb_t1, rtol=2e-6, atol=2e-7) and torch.allclose(emb_s0, emb_s1, rtol=2e-6, atol=2e-7), (
            float((emb_t0 - emb_t1).abs().max()), float((emb_s0 - emb_s1).abs().max()))
    else:
        assert torch.equal(emb_t0, emb_t1) and torch.equal(emb_s0, emb_s1), (float((emb_t0 - emb_t1).abs().max()), float((emb_s0 - emb_s1).abs().max()))
    assert torch.equal(bits0, bits1) and torch.equal(lens0, lens1)
    assert torch.equal(xs0.view(torch.int16), xs1.view(torch.int16))


@pytest.mark.parametrize("cta_pair", [2, 1, 0], ids=["transposed-pair", "pair", "single"])
@pytest.mark.parametrize("B,T", [(5, 200), (3, 97), (37, 41)])
def test_fused_noisy_forward_vs_oracle(B, T, cta_pair, tc_switch):
    """The fused kernel against the oracle's encoder on the oracle's own weak / strong views (supplied noise)."""
    from dad_b200 import ops
    from oracle import dad_oracle as O
    tc_switch(cta_pair=cta_pair)
    st = pretrained_state()
    Ws, bs = st["student_encoder.pre_net.weight"], st["student_encoder.pre_net.bias"]
    g = torch.Generator().manual_seed(5)
    Wt, bt = Ws + 0.01 * torch.randn(256, 768, generator=g), bs + 0.01 * torch.randn(256, generator=g)
    x, pad, aw, as_, hold, (n, rng) = _noisy_views(B, T, 31 + B * 1000 + T, False)
    xh, ph = n["net_input"]["feats"], n["net_input"]["padding_mask"]
    ref_t = O.encoder(O.weak_augment(xh, rng["n_weak"], 0.01), ph, Wt, bt)
    ref_s = O.encoder(O.strong_augment(xh, rng["n_strong"], 0.05, rng["feat_keep_u"], 0.1, rng["t_start"], 0.1), ph, Ws, bs)
    emb_t, emb_s, bits, lens = ops.aug_encoder_fwd(x, pad, Wt.cuda(), bt.cuda(), Ws.cuda(), bs.cuda(), aw, as_)
    assert torch.allclose(emb_t.cpu(), ref_t, rtol=1e-4, atol=2e-5), float((emb_t.cpu() - ref_t).abs().max())
    assert torch.allclose(emb_s.cpu(), ref_s, rtol=1e-4, atol=2e-5), float((emb_s.cpu() - ref_s).abs().max())
    assert torch.equal(lens.cpu(), (~ph).sum(1).float())


@pytest.mark.parametrize("env", [{"DAD_TC_TP_NF": "128"}, {"DAD_TC_TP_TEAMS": "3"}], ids=["256-frame-tiles", "three-converter-teams"])
def test_transposed_kernel_experiment_variants_pass_the_same_parity_tests(env):
    """The two opt-in forms of k_enc_fwd_tp that DESIGN.md 5.1 reports as measured-and-slower (N = 256 frames per tile; 24 converter
    warps) are selected by environment variables read once per process, so they are exercised in a child process: the fused-forward
    parity tests above (bit-identical ReLU bits / operand cache, embeddings to fp32 rounding, oracle) must pass unchanged."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_ops.py"), "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider",
                        "-k", "fused_noisy and transposed and not experiment"], env={**os.environ, **env}, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.parametrize("B,T,padded", [(6, 64, True), (37, 41, True), (64, 200, False), (5, 97, True)])
def test_fused_predict_vs_oracle(B, T, padded):
    """dad_predict: student + teacher logits / arg-max from one read of x == the oracle's encoder + classifier in
    eval mode (I/model.py:225-245): logits 1e-5 (+1e-4 relative), arg-max exact away from ties, so the student / teacher
    disagreement rate of validate (I/train.py:522-564) is exact; the one-network form agrees bit for bit with the two-network one."""
    from dad_b200 import ops
    from oracle import dad_oracle as O
    st = pretrained_state()
    c, n, _ = _inputs(B, T, 900 + B + T, variable=padded)
    x, pad = n["net_input"]["feats"], n["net_input"]["padding_mask"]
    g = torch.Generator().manual_seed(3)
    Ws, bs = st["student_encoder.pre_net.weight"], st["student_encoder.pre_net.bias"]
    Wcs, bcs = st["student_classifier.fc_layer.weight"], st["student_classifier.fc_layer.bias"]
    Wt, bt = Ws + 0.02 * torch.randn(256, 768, generator=g), bs + 0.02 * torch.randn(256, generator=g)
    Wct, bct = Wcs + 0.05 * torch.randn(4, 256, generator=g), bcs + 0.05 * torch.randn(4, generator=g)
    ref_s = O.classifier(O.encoder(x, pad, Ws, bs), None, 0.0, Wcs, bcs)
    ref_t = O.classifier(O.encoder(x, pad, Wt, bt), None, 0.0, Wct, bct)
    cu = lambda t: t.cuda()
    o = ops.predict(cu(x), cu(pad) if padded else None, cu(Ws), cu(bs), cu(Wcs), cu(bcs), cu(Wt), cu(bt), cu(Wct), cu(bct))
    for got, ref in ((o["logits_s"], ref_s), (o["logits_t"], ref_t)):
        assert torch.allclose(got.cpu(), ref, rtol=1e-4, atol=1e-5), float((got.cpu() - ref).abs().max())
    unclear = torch.zeros(B, dtype=torch.bool)
    for pred, ref in ((o["pred_s"], ref_s), (o["pred_t"], ref_t)):
        top2 = ref.topk(2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4
        unclear |= ~clear
        assert torch.equal(pred.cpu().long()[clear], ref.argmax(1)[clear])
    # validate's student / teacher disagreement rate: exact unless a sample sits on a logit tie
    dis_ref = float((ref_s.argmax(1) != ref_t.argmax(1)).float().mean())
    dis_got = float((o["pred_s"] != o["pred_t"]).float().mean())
    assert abs(dis_got - dis_ref) <= float(unclear.sum()) / B + 1e-12
    one = ops.predict(cu(x), cu(pad) if padded else None, cu(Ws), cu(bs), cu(Wcs), cu(bcs))
    assert torch.equal(one["logits_s"], o["logits_s"]) and torch.equal(one["pred_s"], o["pred_s"])


def test_model_predict_uses_the_fused_path_and_matches_the_module_path():
    """SSRLModel.predict / predict_both (drop-in surface, I/model.py:225-245) on the fused tcgen05 pass == the module path
    (torch.ops.dad.encoder_pool + nn.Linear) within the kernels' tolerance; eval mode is set as upstream."""
    from dad_b200 import config as cfg
    from dad_b200.model import SSRLModel
    model = SSRLModel(cfg).cuda()
    model.load_state_dict(pretrained_state())
    with torch.no_grad():
        for p_ in model.teacher_encoder.parameters():
            p_.add_(0.01 * torch.randn_like(p_))
    c, n, _ = _inputs(9, 70, 77)
    x, pad = n["net_input"]["feats"].cuda(), n["net_input"]["padding_mask"].cuda()
    model.train()
    ls = model.predict(x, pad)
    assert not model.training  # I/model.py:237
    lt = model.predict(x, pad, use_teacher=True)
    bs_, bt_ = model.predict_both(x, pad)
    assert torch.allclose(ls, bs_, rtol=1e-5, atol=1e-6) and torch.allclose(lt, bt_, rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        ref_s = model.student_classifier(model.student_encoder(x, pad))
        ref_t = model.teacher_classifier(model.teacher_encoder(x, pad))
    assert torch.allclose(ls, ref_s, rtol=1e-4, atol=1e-5) and torch.allclose(lt, ref_t, rtol=1e-4, atol=1e-5)
    short = model.predict(x[:, :20].contiguous(), pad[:, :20].contiguous())   # T < 32: CUDA-core encoder kernel
    assert short.shape == (9, 4)


def test_philox_noise_field_is_global_under_data_parallel_offsets():
    """the noise of rank r must be the frames [r*B*T, (r+1)*B*T) of ONE global field (not the same field on every
    rank): a batch augmented in two halves with frame offsets equals the batch augmented at once."""
    from dad_b200 import _lib
    import ctypes as C
    B, T = 4, 40
    x = torch.randn(B, T, 768, device="cuda")
    out = torch.empty_like(x)

    def run(xv, ov, off):
        a = _lib.make_aug(1, sigma=1.0, seed=9, stream=_lib.STREAM_WEAK, frame_offset=off)
        _lib.check(_lib.lib().dad_augment(xv.data_ptr(), ov.data_ptr(), C.byref(a), xv.shape[0], T, _lib.stream_ptr()))
    run(x, out, 0)
    halves = torch.empty_like(x)
    run(x[:2].contiguous(), halves[:2], 0)
    run(x[2:].contiguous(), halves[2:], 2 * T)
    assert torch.equal(out, halves)
    same = torch.empty_like(x[:2])
    run(x[2:].contiguous(), same, 0)                      # the defect: rank 1 reusing rank 0's part of the field
    assert not torch.equal(same, halves[2:])


def test_encoder_rejects_bad_arguments():
    from dad_b200 import _lib, ops
    with pytest.raises(_lib.DadError):
        ops.encoder_pool_fwd(torch.zeros(2, 4, 100, device="cuda"), None, torch.zeros(256, 768, device="cuda"), torch.zeros(256, device="cuda"))
    with pytest.raises(_lib.DadError):
        ops.encoder_pool_fwd(torch.zeros(2, 4, 768), None, torch.zeros(256, 768), torch.zeros(256))  # CPU tensors: no fallback
    emb, bits, lens = ops.encoder_pool_fwd(torch.zeros(0, 4, 768, device="cuda"), None, torch.zeros(256, 768, device="cuda"),
                                           torch.zeros(256, device="cuda"))
    assert emb.shape == (0, 256)


def test_ema_matches_reference_rounding():
    from dad_b200 import ops
    from oracle import dad_oracle as O
    g = torch.Generator().manual_seed(3)
    t, s = torch.randn(197892, generator=g), torch.randn(197892, generator=g)
    out = ops.ema_(t.clone().cuda(), s.cuda(), 0.995)
    assert torch.equal(out.cpu(), O.ema_update(t, s, 0.995))


def test_autograd_module_path_matches_oracle():
    """Drop-in module API: SSRLModel.student_encoder(x, pad) is differentiable through torch.ops.dad.encoder_pool."""
    from dad_b200 import config as cfg
    from dad_b200.model import SSRLModel
    from oracle import dad_oracle as O
    st = pretrained_state()
    model = SSRLModel(cfg).cuda()
    model.load_state_dict(st)
    c, n, rng = _inputs(4, 50, 99)
    x, pad = c["net_input"]["feats"], c["net_input"]["padding_mask"]
    emb = model.student_encoder(x.cuda(), pad.cuda())
    emb.square().sum().backward()
    W = st["student_encoder.pre_net.weight"].clone().requires_grad_(True)
    b = st["student_encoder.pre_net.bias"].clone().requires_grad_(True)
    ref = O.encoder(x, pad, W, b)
    ref.square().sum().backward()
    assert torch.allclose(emb.detach().cpu(), ref.detach(), rtol=1e-4, atol=1e-5)
    gW = model.student_encoder.pre_net.weight.grad.cpu()
    assert float((gW - W.grad).abs().max()) <= 1e-4 * float(W.grad.abs().max())
    assert model.teacher_encoder.pre_net.weight.grad is None


# ------------------------------------------------------------------------------------------------ op-level classes (dad_b200.utils)
def test_data_augmentation_class_semantics():
    """utils.DataAugmentation keeps the reference's semantics (I/utils.py:317-375): weak = x + N(0, 0.01^2); strong zeroes one
    shared set of feature columns (no rescale) and one span of int(T*0.1) frames per sample; 2-D and 3-D inputs."""
    from dad_b200.utils import DataAugmentation
    torch.manual_seed(0)
    aug = DataAugmentation()
    x = torch.randn(6, 50, 768, device="cuda")
    w = aug.weak_augment(x)
    assert w.shape == x.shape and abs(float((w - x).std()) - 0.01) < 5e-4
    s = aug.strong_augment(x)
    zero_cols = (s == 0).all(dim=0).all(dim=0)               # feature columns dropped for the whole batch
    assert 30 <= int(zero_cols.sum()) <= 130                 # ~10 % of 768
    live = s[:, :, ~zero_cols]
    zero_rows = (live == 0).all(dim=2)                       # [B,T]
    assert torch.equal(zero_rows.sum(dim=1).cpu(), torch.full((6,), 5))  # int(50 * 0.1) frames per sample
    for b in range(6):                                        # contiguous span
        idx = zero_rows[b].nonzero().flatten()
        assert int(idx[-1] - idx[0]) == 4
    keep_rows = ~zero_rows
    d = (s - x)[:, :, ~zero_cols][keep_rows]
    assert abs(float(d.std()) - 0.05) < 2e-3
    s2 = aug.strong_augment(x[0])
    assert s2.shape == x[0].shape


def test_dacp_manager_matches_oracle_over_steps():
    from dad_b200 import config as cfg
    from dad_b200.utils import DACPManager
    from oracle import dad_oracle as O
    hp = O.HParams.from_cfg(cfg)
    st = O.DACPState.fresh(4)
    mgr = DACPManager(4, cfg.EPOCHS, "cuda")
    anchors = torch.tensor([0.05, 0.3, 0.1, 0.0])
    g = torch.Generator().manual_seed(5)
    for step in range(12):
        B = 64 if step % 3 else 7
        probs = torch.softmax(3.0 * torch.randn(B, 4, generator=g), dim=1)
        if step == 4:
            probs[:, 2] = 0; probs = probs / probs.sum(1, keepdim=True)   # an empty predicted class -> threshold falls back to theta
        s_ref, p_ref = O.certainty_scores(probs, True)
        s_got, p_got = DACPManager.calculate_certainty_scores(probs.cuda())
        assert torch.equal(p_got.cpu(), p_ref) and torch.allclose(s_got.cpu(), s_ref, rtol=2e-6, atol=1e-7)
        epoch = 40 + step
        m_ref, sc_ref, w_ref = O.dacp_mask(st, probs, epoch, anchors, hp)
        m_got, sc_got, w_got = mgr.calculate_mask(probs.cuda(), epoch, anchors.cuda())
        near = (sc_ref - st.ema_thr[p_ref]).abs() < 1e-6        # samples within an ulp of the threshold may legitimately flip
        assert torch.equal(m_got.cpu()[~near], m_ref[~near])
        assert torch.allclose(mgr.ema_thresholds.cpu(), st.ema_thr, rtol=0, atol=1e-6)
        assert torch.allclose(w_got.cpu(), w_ref, rtol=1e-6, atol=1e-7)
        if step % 4 == 3:
            O.dacp_epoch_update(st, hp)
            mgr.update_class_quality_scores_epoch(mgr.batch_scores_per_class)
            assert torch.allclose(mgr.class_quality_scores.cpu(), st.quality, rtol=0, atol=1e-6)


@pytest.mark.parametrize("class_aware", [True, False])
def test_ecda_loss_module_forward_backward_vs_oracle(class_aware):
    from dad_b200.utils import ECDALoss, use_config
    from oracle import dad_oracle as O
    from helpers import cfg_for
    cfg = cfg_for("iemocap", {"USE_CLASS_AWARE_MMD": class_aware})
    use_config(cfg)
    try:
        hp = O.HParams.from_cfg(cfg)
        g = torch.Generator().manual_seed(21)
        B = 48
        mu = torch.randn(4, 256, generator=g)
        yc, yn = torch.randint(0, 4, (B,), generator=g), torch.randint(0, 4, (B,), generator=g)
        clean = (mu[yc] + 0.5 * torch.randn(B, 256, generator=g)).requires_grad_(True)
        noisy = (mu[yn] + 0.7 * torch.randn(B, 256, generator=g)).requires_grad_(True)
        mask = torch.rand(B, generator=g) > 0.3
        scores = torch.rand(B, generator=g)
        w = torch.rand(4, generator=g)
        ref = O.ecda_loss(clean, noisy, yc, yn, mask, scores, w, hp)
        ref.backward()
        c2, n2 = clean.detach().cuda().requires_grad_(True), noisy.detach().cuda().requires_grad_(True)
        got = ECDALoss()(c2, n2, yc.cuda(), yn.cuda(), mask.cuda(), scores.cuda(), w.cuda())
        (2.0 * got).backward()
        assert abs(float(got) - float(ref)) <= 1e-4 * max(abs(float(ref)), 1e-3), (float(got), float(ref))
        for a, b in ((c2.grad, clean.grad), (n2.grad, noisy.grad)):
            assert float((a.cpu() - 2.0 * b).abs().max()) <= 2e-4 * float(b.abs().max()) + 1e-7
    finally:
        from dad_b200 import config
        use_config(config)


def test_device_batch_assembly_equals_host_collator():
    """dad_gather_pad (device-side collate) == the host collator == the reference layout, bit for bit."""
    from dad_b200 import dataload
    fs, _ = dataload.synthetic_feature_sets(5, 7, t_min=3, t_max=41, seed=5)
    idx = list(range(len(fs)))
    host = dataload.UtteranceLoader(fs, idx, 6, shuffle=False)
    dev = dataload.DeviceUtteranceLoader(fs, idx, 6, shuffle=False, device="cuda")
    n = 0
    for hb, db in zip(host, dev):
        assert db["net_input"]["feats"].is_cuda and db["net_input"]["padding_mask"].dtype == torch.bool
        assert torch.equal(db["net_input"]["feats"].cpu(), hb["net_input"]["feats"])
        assert torch.equal(db["net_input"]["padding_mask"].cpu(), hb["net_input"]["padding_mask"])
        assert torch.equal(db["labels"].cpu(), hb["labels"]) and torch.equal(db["id"], hb["id"])
        n += 1
    assert n == len(host) == 6


@pytest.mark.parametrize("B,class_aware", [(700, True), (300, False), (4096, True), (2048, False)])
def test_ecda_multi_tile_vs_fp64_gram_oracle(B, class_aware):
    """Large classes: several Gram row / column blocks, many (P Z') k-blocks and split-K slices per class (the tiny fixtures use
    one tile).  Checked against the oracle's fp64 Gram-form restatement (itself validated against the reference form at small B)."""
    from dad_b200.utils import ECDALoss, use_config
    from oracle import dad_oracle as O
    from helpers import cfg_for
    cfg = cfg_for("iemocap", {"USE_CLASS_AWARE_MMD": class_aware})
    use_config(cfg)
    try:
        hp = O.HParams.from_cfg(cfg)
        g = torch.Generator().manual_seed(77)
        mu = torch.randn(4, 256, generator=g)
        yc, yn = torch.randint(0, 4, (B,), generator=g), torch.randint(0, 4, (B,), generator=g)
        clean = (mu[yc] + 0.5 * torch.randn(B, 256, generator=g))
        noisy = (mu[yn] + 0.7 * torch.randn(B, 256, generator=g))
        mask = torch.rand(B, generator=g) > 0.2
        scores = torch.rand(B, generator=g) * 0.5 + 0.5
        w = torch.rand(4, generator=g)
        c64, n64 = clean.double().requires_grad_(True), noisy.double().requires_grad_(True)
        ref = O.ecda_loss_gram(c64, n64, yc, yn, mask, scores.double(), w.double(), hp)
        ref.backward()
        c2, n2 = clean.cuda().requires_grad_(True), noisy.cuda().requires_grad_(True)
        got = ECDALoss()(c2, n2, yc.cuda(), yn.cuda(), mask.cuda(), scores.cuda(), w.cuda())
        got.backward()
        assert abs(float(got.detach()) - float(ref.detach())) <= 1e-4 * abs(float(ref.detach())), (float(got.detach()), float(ref.detach()))
        for a, b in ((c2.grad, c64.grad), (n2.grad, n64.grad)):
            assert float((a.cpu().double() - b).abs().max()) <= 2e-4 * float(b.abs().max()), float((a.cpu().double() - b).abs().max()) / float(b.abs().max())
    finally:
        from dad_b200 import config
        use_config(config)
