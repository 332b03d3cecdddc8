"""CPU: pin oracle/dad_oracle.py against the golden vectors produced by the unmodified reference
(tests/golden/*.pt, generator oracle/make_golden.py) and, when /root/reference is present, the live reference.

Tolerances: pure element-wise results (augmentation, masks, pseudo-labels) are compared bit-exactly on every step.
Anything behind a CPU GEMM / reduction / libm call is compared to a few fp32 ulps, because the fixtures were written on
one build container's CPU and ATen's SIMD reduction order differs between hosts (AVX2 vs AVX-512).  Those 1-ulp
differences are amplified by Adam (lr 5e-3, update g/sqrt(v) is O(lr) however small g is), so a free-running trajectory
of the UNMODIFIED reference drifts away from its own fixture on another host: measured reference(AVX-512 host) vs
fixture(AVX2 host): losses 2e-6 after 10 steps and 3e-5 after 100 steps, weights 3.1e-5 after 10 steps (step_emodb);
the oracle on the AVX-512 host sits on the live reference of that host to 1e-6 (lockstep tests at the bottom).  Hence:
  * steps 0..9 of every case: LOSS_RTOL, the tight cross-host pin (the trajectory has not had time to drift);
  * steps >= 10 (the 100-step case only): DRIFT_RTOL, a basin check, with masks / pseudo-labels still bit-exact;
  * final weights: W_ATOL_10 after 10 steps, W_ATOL_100 after 100;
  * the same-host lockstep tests (live reference, every case, every step) keep the 1e-6 bound wherever
    /root/reference is mounted (the container the CPU suite is graded in)."""
LOSS_RTOL = 2e-5    # north_star loss tolerance is 1e-4
DRIFT_RTOL = 1e-2   # free-running cross-host drift bound for steps >= 10 (measured up to 2.4e-3 at step 85)
W_ATOL_10 = 1e-4    # cross-host, 10 Adam steps (measured 3.1e-5, identical for oracle and live reference)
W_ATOL_100 = 5e-3   # cross-host, 100 Adam steps (measured 6.1e-4)
import pytest
import torch

from helpers import load_golden, make_oracle, pretrained_state, step_inputs
from dad_b200 import synth
from oracle import dad_oracle as O
from oracle import ref_harness

CASES = ["step_iemocap_warmup", "step_iemocap_noentropy_globalmmd", "step_iemocap_fixedthr", "step_casia_default",
         "step_casia_full", "step_emodb", "step_iemocap",
         "step_emodb_b32_t200", "step_iemocap_b64_t200"]  # the last two: BASELINE.json configs[0] / configs[1] at their exact shapes


def run_case(name, max_steps=None):
    fix = load_golden(name)
    meta = fix["meta"]
    orc, cfg, hp = make_oracle(meta["dataset"], meta["overrides"])
    k = 0
    for epoch in meta["epochs"]:
        orc.lr = O.cosine_lr(hp.LEARNING_RATE, epoch, hp.EPOCHS)
        for _ in range(meta["steps_per_epoch"]):
            c, n, rng = step_inputs(meta, k, cfg)
            if k == 0:
                cs = fix["input_checksums"][0]
                assert abs(float(c["net_input"]["feats"].double().sum()) - cs["x_clean"][0]) < 1e-6, "synthetic input RNG drift"
                assert rng["t_start"].tolist() == cs["t_start"]
            out = orc.step(c, n, epoch, rng)
            ref = fix["steps"][k]
            tol = LOSS_RTOL if k < 10 else DRIFT_RTOL
            for key in ("total_loss", "supervised_ce_loss", "consistency_loss", "ecda_loss", "grad_norm"):
                assert abs(out[key] - ref[key]) <= tol * max(1.0, abs(ref[key])), (name, k, key, out[key], ref[key])
            if fix["mask"] is not None and epoch >= hp.WARMUP_EPOCHS:
                assert torch.equal(orc.last["mask"].to(torch.uint8), fix["mask"][k]), (name, k)
                assert torch.equal(orc.last["pred"].to(torch.uint8), fix["pred"][k])
            k += 1
            if max_steps and k >= max_steps:
                return
        orc.end_epoch(epoch)
    sd = orc.state_dict()
    stride = fix["final_state_stride"]
    for key, ref in fix["final_state"].items():
        got = sd[key] if (stride == 1 or sd[key].numel() < 2048) else sd[key].flatten()[::stride]
        assert torch.allclose(got, ref, rtol=0, atol=W_ATOL_10 if k <= 10 else W_ATOL_100), (name, key, (got - ref).abs().max())
    assert torch.allclose(orc.dacp.ema_thr, fix["dacp_ema_thr"], atol=1e-6 if k <= 10 else 1e-4)
    assert torch.allclose(orc.dacp.quality, fix["dacp_quality"], atol=1e-6 if k <= 10 else 1e-4)


@pytest.mark.parametrize("name", CASES)
def test_oracle_step_matches_reference_golden(name):
    run_case(name)


def test_oracle_ops_match_reference_golden():
    g = load_golden("ops_small")
    cfg_drop, B, T = 0.1, g["B"], g["T"]
    c, n = synth.make_batches(B, T, g["seed_batches"], variable_len=True)
    rng = synth.make_rng_inputs(B, T, g["seed_rng"])
    xn = n["net_input"]["feats"]
    assert torch.equal(O.weak_augment(xn, rng["n_weak"], 0.01), g["weak"])
    assert torch.equal(O.strong_augment(xn, rng["n_strong"], 0.05, rng["feat_keep_u"], cfg_drop, rng["t_start"], 0.1), g["strong"])
    assert torch.equal(O.strong_augment(xn[0], rng["n_strong"][0], 0.05, rng["feat_keep_u"], cfg_drop, rng["t_start"][:1], 0.1), g["strong2d"])
    st = pretrained_state()
    W, b = st["student_encoder.pre_net.weight"], st["student_encoder.pre_net.bias"]
    assert torch.allclose(O.encoder(c["net_input"]["feats"], c["net_input"]["padding_mask"], W, b), g["emb_pad"], rtol=0, atol=1e-6)
    assert torch.allclose(O.encoder(c["net_input"]["feats"], None, W, b), g["emb_nopad"], rtol=0, atol=1e-6)
    # DACP incl. an empty class, three consecutive calls + epoch update
    hp = O.HParams()
    d = g["dacp"]
    stt = O.DACPState.fresh(4)
    stt.quality = d["quality0"].clone()
    for o in d["out"]:
        m, s, w = O.dacp_mask(stt, d["P"], o["epoch"], d["anchors"], hp)
        assert torch.equal(m, o["mask"])
        assert torch.allclose(s, o["scores"], rtol=0, atol=1e-6) and torch.allclose(w, o["w"], rtol=0, atol=1e-6)
        assert torch.allclose(stt.ema_thr, o["ema_thr"], rtol=0, atol=1e-6)
    O.dacp_epoch_update(stt, hp)
    assert torch.allclose(stt.quality, d["quality_after_epoch"], atol=1e-7)
    # ECDA forward + autograd grads, class-aware and global; and the fp64 Gram-form restatement
    e = g["ecda"]
    for aware, key in ((True, "aware"), (False, "global")):
        hp2 = O.HParams(USE_CLASS_AWARE_MMD=aware)
        fc, fn = e["fc"].clone().requires_grad_(True), e["fn"].clone().requires_grad_(True)
        loss = O.ecda_loss(fc, fn, e["yc"], e["yn"], e["mask"], e["scores"], e["wce"], hp2)
        loss.backward()
        assert torch.allclose(loss.detach(), e["out"][key]["loss"], rtol=1e-5, atol=0)
        assert torch.allclose(fc.grad, e["out"][key]["g_clean"], rtol=1e-4, atol=1e-7)
        assert torch.allclose(fn.grad, e["out"][key]["g_noisy"], rtol=1e-4, atol=1e-7)
        l64 = O.ecda_loss_gram(e["fc"].double(), e["fn"].double(), e["yc"], e["yn"], e["mask"], e["scores"].double(), e["wce"].double(), hp2)
        assert abs(float(l64) - float(e["out"][key]["loss"])) < 2e-5 * abs(float(e["out"][key]["loss"]))


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_matches_live_reference():
    """Bit-level lockstep with the imported reference on fresh seeds (beyond what the fixtures hold)."""
    ov = {"WARMUP_EPOCHS": 1, "ECDA_START_EPOCH": 1, "WEIGHT_RAMP_EPOCHS": 2}
    ref = ref_harness.RefStepper("iemocap", overrides=ov, state=pretrained_state())
    orc, cfg, hp = make_oracle("iemocap", ov)
    k = 0
    for epoch in (0, 1, 2, 3):
        for _ in range(3):
            c, n = synth.make_batches(12, 24, 31000 + k, variable_len=True)
            rng = synth.make_rng_inputs(12, 24, 32000 + k)
            a, b = ref.step(c, n, epoch, rng), orc.step(c, n, epoch, rng)
            for key in b:
                assert a[key] == pytest.approx(b[key], rel=1e-6, abs=1e-7), (epoch, key)
            k += 1
        if epoch >= hp.WARMUP_EPOCHS:
            ref.tr.dacp_manager.update_class_quality_scores_epoch(ref.tr.dacp_manager.batch_scores_per_class)
        orc.end_epoch(epoch)
    sd, so = ref.tr.model.state_dict(), orc.state_dict()
    for key in sd:
        assert torch.allclose(sd[key], so[key], atol=1e-7), key


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("name", CASES)
def test_oracle_lockstep_with_live_reference_on_golden_cases(name):
    """Same-host pin for every fixture case: the imported reference and the oracle run side by side over the case's
    whole trajectory (100 steps for step_iemocap) from the same state on the same inputs; every reported scalar
    agrees to 1e-6 on every step and every weight to 1e-6 at the end (free-running, no teacher forcing)."""
    fix = load_golden(name)
    meta = fix["meta"]
    ref = ref_harness.RefStepper(meta["dataset"], overrides=meta["overrides"], state=pretrained_state())
    orc, cfg, hp = make_oracle(meta["dataset"], meta["overrides"])
    k = 0
    for epoch in meta["epochs"]:
        orc.lr = O.cosine_lr(hp.LEARNING_RATE, epoch, hp.EPOCHS)
        for g in ref.tr.optimizer.param_groups:
            g["lr"] = orc.lr
        for _ in range(meta["steps_per_epoch"]):
            c, n, rng = step_inputs(meta, k, cfg)
            a, b = ref.step(c, n, epoch, rng), orc.step(c, n, epoch, rng)
            for key in b:
                assert a[key] == pytest.approx(b[key], rel=1e-6, abs=1e-7), (name, k, key)
            k += 1
        if epoch >= hp.WARMUP_EPOCHS and hasattr(ref.tr, "dacp_manager") and ref.tr.dacp_manager is not None:
            ref.tr.dacp_manager.update_class_quality_scores_epoch(ref.tr.dacp_manager.batch_scores_per_class)
        orc.end_epoch(epoch)
    sd, so = ref.tr.model.state_dict(), orc.state_dict()
    for key in sd:
        assert torch.allclose(sd[key], so[key], rtol=0, atol=1e-6), (name, key, (sd[key] - so[key]).abs().max())
