"""GPU (-m gpu): parity at BASELINE.json's FULL sizes, through the C ABI.

configs[3] (global batch 4096, T=200 on one GPU) and configs[4] (batch 512 per GPU, T=1000) are far beyond what the CPU
oracle finishes in seconds (322 GFLOP per encoder call; the reference's ECDA would materialise [n,n,256] differences), so
the full-size checks are built from three pieces that together cover the whole step:

  1. encoder forward / backward on the full batch against an fp64 torch restatement evaluated ON THE DEVICE in chunks
     (same augmentation arithmetic: fp32 `x + n*sigma`, feature mask, zeroed span; then fp64 GEMMs), which is itself tied to
     the CPU oracle on utterances sampled from the full batch (first, last, random);
  2. everything downstream of the three encoder calls -- heads, CE, certainty scores, DACP quantiles / thresholds / mask, KD,
     ECDA (fp64 Gram-form oracle restatement, validated against the reference form at small B in the CPU suite), and the
     gradient that flows back into the embeddings -- against the oracle's own `forward_losses` + autograd fed with the engine's
     embeddings and dropout masks;
  3. (tests/test_gpu_ops.py::test_ecda_multi_tile_vs_fp64_gram_oracle at B=4096) the ECDA op alone.

Tolerances are stated at each assert; measured values are printed (run with -s to see them)."""
import pytest
import torch

from helpers import make_oracle, pretrained_state, push_oracle_state
from dad_b200 import synth
from oracle import dad_oracle as O

pytestmark = pytest.mark.gpu
SIGMA, FEAT_RATE, MASK_RATIO = 0.05, 0.1, 0.1


def _device_inputs(B, T, seed, variable_len=True, snr_db=None):
    """Class-structured synthetic features generated on the device (SURVEY.md 8d formulas): x = mu[y] + randn."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(seed)
    mu = synth.class_means().to(dev)
    y = torch.randint(0, 4, (B,), generator=g, device=dev)
    x = mu[y][:, None, :] + torch.randn(B, T, 768, generator=g, device=dev)
    if snr_db is not None:  # noisy domain: white noise at the given SNR per utterance
        P = (x * x).mean(dim=(1, 2))
        x = x + torch.sqrt(P / (10.0 ** (snr_db / 10.0)))[:, None, None] * torch.randn(B, T, 768, generator=g, device=dev)
    if variable_len:
        lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=dev)
    else:
        lens = torch.full((B,), T, device=dev)
    pad = torch.arange(T, device=dev)[None, :] >= lens[:, None]
    return x.contiguous(), pad.contiguous(), y, g


def _augment_on_device(x, noise, keep, ts, L):
    """I/utils.py:333-375 in torch on the device: fp32 x + n*sigma (two roundings), shared feature mask, zeroed span."""
    T = x.shape[1]
    xa = x + noise * SIGMA
    xa = xa * keep
    if L > 0:
        ar = torch.arange(T, device=x.device)
        span = (ar[None, :] >= ts[:, None]) & (ar[None, :] < (ts[:, None] + L))
        xa = xa.masked_fill(span[:, :, None], 0.0)
    return xa


def _device_reference(x, pad, aug_in, W, b, g_emb, bits, chunk=256):
    """fp64 restatement of aug + encoder forward, and of dW / db given the engine's 1-bit ReLU decisions `bits` (which are
    themselves compared with the fp64 pre-activations outside the tie zone).  -> emb, dW, db, n_bad_bits, n_valid"""
    B, T, _ = x.shape
    dev = x.device
    W64, b64 = W.double(), b.double()
    emb = torch.empty(B, 256, dtype=torch.float64, device=dev)
    dW = torch.zeros(256, 768, dtype=torch.float64, device=dev)
    db = torch.zeros(256, dtype=torch.float64, device=dev)
    shifts = torch.arange(32, device=dev, dtype=torch.int32)
    bad = 0
    for c0 in range(0, B, chunk):
        c1 = min(B, c0 + chunk)
        xa = x[c0:c1]
        if aug_in is not None:
            noise, keep, ts, L = aug_in
            xa = _augment_on_device(xa, noise[c0:c1], keep, ts[c0:c1], L)
        valid = (~pad[c0:c1]).double()
        lens = valid.sum(1).clamp(min=1.0)
        xa64 = xa.double().reshape(-1, 768)
        pre = xa64 @ W64.t() + b64
        h = pre.clamp(min=0).reshape(c1 - c0, T, 256) * valid[:, :, None]
        emb[c0:c1] = h.sum(1) / lens[:, None]
        word = bits[c0 * T:c1 * T]
        pos = ((word[:, :, None] >> shifts) & 1).reshape(-1, 256).bool()
        sure = (pre.abs() > 1e-4) & valid.reshape(-1, 1).bool()
        bad += int(((pos != (pre > 0)) & sure).sum())
        coef = (g_emb[c0:c1].double() / lens[:, None])[:, None, :] * valid[:, :, None]
        dpre = pos.double() * coef.reshape(-1, 256)
        dW += dpre.t() @ xa64
        db += dpre.sum(0)
    return emb, dW, db, bad


@pytest.mark.parametrize("B,T,aug,mode", [(4096, 200, 2, "tc"), (4096, 200, 0, "tc"), (512, 1000, 2, "tc"), (4096, 200, 2, "fp32")])
def test_encoder_full_size_forward_backward(B, T, aug, mode):
    """BASELINE configs[3] / configs[4] shapes, variable-length batch, strong view with SUPPLIED noise (or the clean view):
    every embedding, every ReLU decision outside the tie zone, dW and db of the full batch."""
    from dad_b200 import _lib, ops
    gm = {"fp32": _lib.GEMM_FP32, "tc": _lib.GEMM_TC}[mode]
    x, pad, _, g = _device_inputs(B, T, 2024 + B + T)
    st = pretrained_state()
    W0, b0 = st["student_encoder.pre_net.weight"], st["student_encoder.pre_net.bias"]
    W, b = W0.cuda(), b0.cuda()
    L = int(T * MASK_RATIO)
    a, aug_in = None, None
    if aug == 2:
        noise = torch.randn(B, T, 768, generator=g, device="cuda")
        keep = (torch.rand(768, generator=g, device="cuda") > FEAT_RATE).float()
        ts = torch.randint(0, T - L + 1, (B,), generator=g, device="cuda").to(torch.int32)
        a = _lib.make_aug(2, noise=noise, sigma=SIGMA, feat_keep=keep, t_start=ts, mask_len=L)
        aug_in = (noise, keep, ts, L)
    g_emb = torch.randn(B, 256, generator=g, device="cuda") * 1e-3
    emb, bits, lens = ops.encoder_pool_fwd(x, pad, W, b, aug=a, gemm_mode=gm)
    dW, db = ops.encoder_pool_bwd(x, pad, bits, g_emb, lens, aug=a, gemm_mode=gm)
    torch.cuda.synchronize()
    emb_ref, dW_ref, db_ref, bad = _device_reference(x, pad, aug_in, W, b, g_emb, bits)
    assert torch.equal(lens, (~pad).sum(1).float())
    # embeddings: same bound as the small-size oracle tests (fp32 kernel 2e-5 / 2e-6, f16x3 tensor-core kernel 1e-4 / 2e-5)
    rtol, atol = (2e-5, 2e-6) if mode == "fp32" else (1e-4, 2e-5)
    err = (emb.double() - emb_ref).abs()
    print(f"[full-size {mode} B={B} T={T} aug={aug}] emb max abs err {float(err.max()):.2e} (|emb| max {float(emb_ref.abs().max()):.2e}), ReLU bits beyond tie zone: {bad}")
    assert bool((err <= atol + rtol * emb_ref.abs()).all()), float(err.max())
    assert bad == 0
    # dW / db of the whole batch: fp32 accumulation of 16.7k-frame split-K partials (round-toward-zero in TMEM on the tcgen05 path)
    tolW = 1e-4 if mode == "fp32" else 2e-4  # measured on B200: 3.5e-5 .. 4.0e-5 (tcgen05 path)
    eW = float((dW.double() - dW_ref).abs().max()) / float(dW_ref.abs().max())
    eb = float((db.double() - db_ref).abs().max()) / float(db_ref.abs().max())
    print(f"[full-size {mode} B={B} T={T} aug={aug}] dW rel-to-max err {eW:.2e}, db {eb:.2e}")
    assert eW <= tolW and eb <= tolW, (eW, eb)
    # tie the device restatement to the CPU oracle on sampled utterances of this very batch
    gi = torch.Generator().manual_seed(5)
    idx = torch.cat([torch.tensor([0, 1, B // 2, B - 1]), torch.randint(0, B, (12,), generator=gi)])
    xs, ps = x[idx.cuda()].cpu(), pad[idx.cuda()].cpu()
    if aug == 2:
        xs = O.strong_augment(xs, noise[idx.cuda()].cpu(), SIGMA, keep.cpu(), FEAT_RATE, ts[idx.cuda()].cpu().long(), MASK_RATIO)
    emb_orc = O.encoder(xs, ps, W0, b0)
    got = emb[idx.cuda()].cpu()
    assert torch.allclose(got, emb_orc, rtol=rtol, atol=atol), float((got - emb_orc).abs().max())
    assert torch.allclose(emb_ref[idx.cuda()].cpu().float(), emb_orc, rtol=2e-5, atol=5e-6)


def test_full_size_step_downstream_of_encoders_vs_oracle():
    """One full DAD step at global batch 4096, T=200, epoch 60 (full loss stack, in-kernel Philox noise).  The oracle's own
    forward_losses (I/train.py:397-471) + autograd is run on the ENGINE'S embeddings and dropout masks (its encoder calls are
    replaced by the engine's three [4096,256] outputs; ECDA by the fp64 Gram restatement): CE, KD, ECDA, total within the
    north-star 1e-4, DACP thresholds 1e-6, mask / pseudo-labels exact, dL/d emb (what the dW kernels consume) 2e-4 of max."""
    from dad_b200 import _lib
    from dad_b200.engine import DADStepEngine
    B, T, epoch = 4096, 200, 60
    orc, cfg, hp = make_oracle("iemocap")
    orc.lr = O.cosine_lr(hp.LEARNING_RATE, epoch, hp.EPOCHS)
    eng = DADStepEngine(cfg, B, T, gemm_mode=_lib.GEMM_TC, use_graph=False, noise_mode="philox", seed=42)
    push_oracle_state(eng, orc)
    xc, pc, yc, _ = _device_inputs(B, T, 31, variable_len=True)
    xn, pn, yn, _ = _device_inputs(B, T, 32, variable_len=True, snr_db=10.0)
    eng.set_inputs({"net_input": {"feats": xc, "padding_mask": pc}, "labels": yc},
                   {"net_input": {"feats": xn, "padding_mask": pn}, "labels": yn})
    eng.step(epoch)
    torch.cuda.synchronize()
    got = eng.losses()
    v = _lib.pack_global_views(eng.pack, B, 1)
    emb_c = v["emb_clean"][0].detach().cpu().clone().requires_grad_(True)
    emb_s = v["emb_strong"][0].detach().cpu().clone().requires_grad_(True)
    emb_t = eng.emb_teacher.detach().cpu().clone()
    queue = [emb_c, emb_t, emb_s]  # call order of I/train.py:399, 409, 439
    dummy = torch.zeros(B, 1, 768)
    nopad = torch.zeros(B, 1, dtype=torch.bool)
    c = {"net_input": {"feats": dummy, "padding_mask": nopad}, "labels": yc.cpu()}
    n = {"net_input": {"feats": dummy, "padding_mask": nopad}, "labels": None}
    rng = {"drop_clean": eng.keep_clean.cpu().clone(), "drop_strong": eng.keep_strong.cpu().clone(), "n_weak": dummy, "n_strong": dummy,
           "feat_keep_u": torch.ones(768), "t_start": torch.zeros(B, dtype=torch.int64)}
    real_enc, real_ecda = O.encoder, O.ecda_loss
    O.encoder = lambda x, pad, W, b: queue.pop(0)
    O.ecda_loss = lambda cl, nz, y1, y2, m, s, w, h: O.ecda_loss_gram(cl.double(), nz.double(), y1, y2, m, s.double(), w.double(), h)
    try:
        ref = orc.forward_losses(c, n, epoch, rng)
        ref["total_loss"].backward()
    finally:
        O.encoder, O.ecda_loss = real_enc, real_ecda
    assert not queue
    refv = {k: float(x.detach()) for k, x in ref.items()}
    flips = int((eng.mask.cpu().bool() != orc.last["mask"]).sum())
    nmask = int(orc.last["mask"].sum())
    print(f"[full-size step] masked-in {nmask}/{B}, mask flips {flips}, losses engine {got} oracle {refv}")
    assert nmask >= 8 and refv["ecda_loss"] != 0.0 and refv["consistency_loss"] > 0.0  # every branch of the stack ran
    assert flips == 0
    assert torch.equal(v["pred"][0].cpu().long(), orc.last["pred"])
    assert float((eng.ema_thr.cpu() - orc.dacp.ema_thr).abs().max()) <= 1e-6
    for key in ("supervised_ce_loss", "consistency_loss", "ecda_loss", "total_loss"):
        r = refv[key]
        assert abs(got[key] - r) <= 1e-4 * max(abs(r), 1e-3), (key, got[key], r)
    for name, g_eng, leaf in (("clean", eng.g_clean, emb_c), ("strong", eng.g_strong, emb_s)):
        d = float((g_eng.cpu().double() - leaf.grad.double()).abs().max()) / float(leaf.grad.abs().max())
        print(f"[full-size step] dL/d emb_{name}: rel-to-max err {d:.2e}")
        assert d <= 2e-4, (name, d)


def test_full_size_step_hard_noisy_domain_partial_mask_through_adam():
    """at the headline size the easy noisy domain masks every sample in, which makes "mask exact" and the KD
    denominator trivial.  Here the noisy domain is at -12 dB and the DACP state sits inside the score distribution: thresholds near
    the per-class 45 % score quantile for classes 0-2 and an anchor above every score for class 3, so that roughly half of the
    batch is masked in and one class is INFEASIBLE for ECDA (no masked noisy member).  Checked: the partial mask and the
    pseudo-labels exactly, thresholds 1e-6, losses 1e-4, dL/d emb, the classifier gradients against the oracle's autograd, and --
    feeding the oracle's clip / Adam / EMA (I/train.py:487-492) the ENGINE'S flat gradient -- the updated student and teacher
    weights at 1e-5 / 1e-6 on all 197 892 elements (same gradient in, so no conditioning caveat)."""
    from dad_b200 import _lib
    from dad_b200.engine import DADStepEngine
    B, T, epoch = 4096, 200, 60
    orc, cfg, hp = make_oracle("iemocap")
    orc.lr = O.cosine_lr(hp.LEARNING_RATE, epoch, hp.EPOCHS)
    orc.anchors = torch.tensor([0.05, 0.05, 0.05, 5.0])
    orc.dacp.ema_thr = torch.tensor([0.898, 0.939, 0.860, 5.0])
    orc.nstep = 7  # mid-run Adam state: bias corrections away from their first-step special case
    g0 = torch.Generator().manual_seed(77)
    for k in O.PARAM_KEYS:
        orc.m[k] = 1e-4 * torch.randn(orc.p[k].shape, generator=g0)
        orc.v[k] = 1e-7 * torch.rand(orc.p[k].shape, generator=g0) + 1e-9
    eng = DADStepEngine(cfg, B, T, gemm_mode=_lib.GEMM_TC, use_graph=False, noise_mode="philox", seed=43)
    push_oracle_state(eng, orc)
    xc, pc, yc, _ = _device_inputs(B, T, 41, variable_len=True)
    xn, pn, yn, _ = _device_inputs(B, T, 42, variable_len=True, snr_db=-12.0)
    eng.set_inputs({"net_input": {"feats": xc, "padding_mask": pc}, "labels": yc},
                   {"net_input": {"feats": xn, "padding_mask": pn}, "labels": yn})
    p_before, t_before = eng.p.cpu().clone(), eng.t.cpu().clone()
    eng.step(epoch)
    torch.cuda.synchronize()
    got = eng.losses()
    v = _lib.pack_global_views(eng.pack, B, 1)
    emb_c = v["emb_clean"][0].detach().cpu().clone().requires_grad_(True)
    emb_s = v["emb_strong"][0].detach().cpu().clone().requires_grad_(True)
    emb_t = eng.emb_teacher.detach().cpu().clone()
    queue = [emb_c, emb_t, emb_s]
    dummy, nopad = torch.zeros(B, 1, 768), torch.zeros(B, 1, dtype=torch.bool)
    c = {"net_input": {"feats": dummy, "padding_mask": nopad}, "labels": yc.cpu()}
    n = {"net_input": {"feats": dummy, "padding_mask": nopad}, "labels": None}
    rng = {"drop_clean": eng.keep_clean.cpu().clone(), "drop_strong": eng.keep_strong.cpu().clone(), "n_weak": dummy, "n_strong": dummy,
           "feat_keep_u": torch.ones(768), "t_start": torch.zeros(B, dtype=torch.int64)}
    real_enc, real_ecda = O.encoder, O.ecda_loss
    O.encoder = lambda x, pad, W, b: queue.pop(0)
    O.ecda_loss = lambda cl, nz, y1, y2, m, s, w, h: O.ecda_loss_gram(cl.double(), nz.double(), y1, y2, m, s.double(), w.double(), h)
    try:
        for p_ in orc.p.values():
            p_.grad = None
        ref = orc.forward_losses(c, n, epoch, rng)
        ref["total_loss"].backward()
    finally:
        O.encoder, O.ecda_loss = real_enc, real_ecda
    refv = {k: float(x.detach()) for k, x in ref.items()}
    mask_o, pred_o = orc.last["mask"], orc.last["pred"]
    nmask = int(mask_o.sum())
    per_class = [int((mask_o & (pred_o == k)).sum()) for k in range(4)]
    flips = int((eng.mask.cpu().bool() != mask_o).sum())
    print(f"[full-size hard step] masked-in {nmask}/{B} per pseudo-class {per_class}, mask flips {flips}, losses engine {got} oracle {refv}")
    assert 0.25 * B <= nmask <= 0.85 * B, nmask                 # a PARTIAL mask
    assert per_class[3] == 0 and min(per_class[:3]) >= 2        # class 3 infeasible for ECDA, the others feasible
    assert flips == 0 and torch.equal(v["pred"][0].cpu().long(), pred_o)
    assert float((eng.ema_thr.cpu() - orc.dacp.ema_thr).abs().max()) <= 1e-6
    for key in ("supervised_ce_loss", "consistency_loss", "ecda_loss", "total_loss"):
        r = refv[key]
        assert abs(got[key] - r) <= 1e-4 * max(abs(r), 1e-3), (key, got[key], r)
    for name, g_eng, leaf in (("clean", eng.g_clean, emb_c), ("strong", eng.g_strong, emb_s)):
        d = float((g_eng.cpu().double() - leaf.grad.double()).abs().max()) / float(leaf.grad.abs().max())
        print(f"[full-size hard step] dL/d emb_{name}: rel-to-max err {d:.2e}")
        assert d <= 2e-4, (name, d)
    # classifier gradients (autograd of the oracle through its own heads) vs the engine's flat gradient
    g_eng = eng.g_ext[:_lib.NPARAM].cpu()
    for key, a0, a1 in (("student_classifier.fc_layer.weight", _lib.OFF_WC, _lib.OFF_BC), ("student_classifier.fc_layer.bias", _lib.OFF_BC, _lib.NPARAM)):
        go = orc.p[key].grad.reshape(-1)
        d = float((g_eng[a0:a1] - go).abs().max()) / float(go.abs().max())
        print(f"[full-size hard step] d{key}: rel-to-max err {d:.2e}")
        assert d <= 2e-4, (key, d)
    # optimiser tail at full size: oracle clip + Adam + EMA on the engine's gradient == the engine's updated weights
    grads = [g_eng[a:b].clone().view(shape) for (a, b, shape) in ((_lib.OFF_W, _lib.OFF_B, (256, 768)), (_lib.OFF_B, _lib.OFF_WC, (256,)),
                                                                  (_lib.OFF_WC, _lib.OFF_BC, (4, 256)), (_lib.OFF_BC, _lib.NPARAM, (4,)))]
    O.clip_grad_norm(grads, hp.MAX_GRAD_NORM)
    orc.nstep += 1
    with torch.no_grad():
        for k, g_ in zip(O.PARAM_KEYS, grads):
            O.adam_step(orc.p[k], g_, orc.m[k], orc.v[k], orc.nstep, orc.lr, hp.WEIGHT_DECAY)
        for ks, kt in zip(O.PARAM_KEYS, O.TEACHER_KEYS):
            orc.t[kt] = O.ema_update(orc.t[kt], orc.p[ks].detach(), hp.EMA_MOMENTUM)
    p_orc = torch.cat([orc.p[k].detach().reshape(-1) for k in O.PARAM_KEYS])
    t_orc = torch.cat([orc.t[k].reshape(-1) for k in O.TEACHER_KEYS])
    dp_, dt_ = float((eng.p.cpu() - p_orc).abs().max()), float((eng.t.cpu() - t_orc).abs().max())
    moved = float((eng.p.cpu() - p_before).abs().max())
    print(f"[full-size hard step] after clip+Adam+EMA: student max abs diff {dp_:.2e} (update size {moved:.2e}), teacher {dt_:.2e}")
    assert moved > 1e-5 and float((eng.t.cpu() - t_before).abs().max()) > 0
    assert dp_ <= 1e-5 and dt_ <= 1e-6, (dp_, dt_)
