"""Generate tests/golden/*.pt by EXECUTING THE UNMODIFIED REFERENCE (via oracle/ref_harness.py).

Run in the build container only:  python -m oracle.make_golden
The fixtures pin oracle/dad_oracle.py (tests/test_oracle_golden.py) and the CUDA path (tests/test_*gpu*.py).
Inputs are regenerated from seeds by dad_b200.synth; input checksums are stored to detect RNG drift.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dad_b200  # noqa: E402,F401
from dad_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STUDENT = ("student_encoder.pre_net.weight", "student_encoder.pre_net.bias",
           "student_classifier.fc_layer.weight", "student_classifier.fc_layer.bias")


def checksum(t):
    return float(t.double().sum()), float(t.double().abs().sum())


def full_state(student):
    s = {k: v.clone() for k, v in student.items()}
    s.update({k.replace("student_", "teacher_"): v.clone() for k, v in student.items()})
    return s


def pretrain():
    """30 CE-only steps (epoch 0 < WARMUP) from the default init -> 'pretrained-like' student."""
    ref = ref_harness.RefStepper("iemocap", state=synth.init_state(42))
    for i in range(30):
        c, n = synth.make_batches(32, 40, 100 + i)
        ref.step(c, n, 0, synth.make_rng_inputs(32, 40, 200 + i))
    sd = ref.tr.model.state_dict()
    return {k: sd[k].detach().clone() for k in STUDENT}


def run_case(name, dataset, student, overrides, B, T, epochs, steps_per_epoch, seed, variable_len=True, snr_db=10.0,
             full_final=False):
    ref = ref_harness.RefStepper(dataset, overrides=overrides, state=full_state(student))
    cfg, tr = ref.cfg, ref.tr
    rec = {"mask": [], "pred": [], "scores": [], "probs": []}

    def hook(_m, _i, logits):
        rec["_probs"] = torch.softmax(logits.detach(), dim=1)

    tr.model.teacher_classifier.register_forward_hook(hook)
    if hasattr(tr.dacp_manager, "calculate_mask"):
        orig = tr.dacp_manager.calculate_mask

        def wrapped(p, e, a):
            m, s, w = orig(p, e, a)
            rec["_dacp"] = (m.clone(), s.clone())
            return m, s, w

        tr.dacp_manager.calculate_mask = wrapped
    steps, sums = [], []
    k = 0
    for epoch in epochs:
        lr = 0.5 * cfg.LEARNING_RATE * (1 + __import__("math").cos(__import__("math").pi * epoch / cfg.EPOCHS))
        for g in tr.optimizer.param_groups:
            g["lr"] = lr
        for _ in range(steps_per_epoch):
            c, n = synth.make_batches(B, T, seed + k, snr_db=snr_db, variable_len=variable_len)
            rng = synth.make_rng_inputs(B, T, seed + 100000 + k, drop_p=cfg.DROPOUT_RATE, mask_ratio=cfg.TEMPORAL_MASK_RATIO)
            rec.pop("_dacp", None)
            rec.pop("_probs", None)
            out = ref.step(c, n, epoch, rng)
            out["epoch"], out["lr"] = epoch, lr
            steps.append(out)
            if "_probs" in rec:
                p = rec["_probs"]
                if "_dacp" in rec:
                    m, s = rec["_dacp"]
                else:
                    s = p.max(1).values
                    m = s >= cfg.FIXED_CONFIDENCE_THRESHOLD
                rec["mask"].append(m.to(torch.uint8))
                rec["scores"].append(s)
                rec["pred"].append(p.argmax(1).to(torch.uint8))
                rec["probs"].append(p)
            if k in (0, len(epochs) * steps_per_epoch - 1):
                sums.append({"step": k, "x_clean": checksum(c["net_input"]["feats"]), "x_noisy": checksum(n["net_input"]["feats"]),
                             "n_strong": checksum(rng["n_strong"]), "t_start": rng["t_start"].tolist()})
            k += 1
        if epoch >= cfg.WARMUP_EPOCHS:
            tr.dacp_manager.update_class_quality_scores_epoch(tr.dacp_manager.batch_scores_per_class)
    sd = tr.model.state_dict()
    fix = {
        "meta": {"name": name, "dataset": dataset, "overrides": overrides, "B": B, "T": T, "epochs": list(epochs),
                 "steps_per_epoch": steps_per_epoch, "seed": seed, "variable_len": variable_len, "snr_db": snr_db,
                 "torch": torch.__version__},
        "steps": steps,
        "mask": torch.stack(rec["mask"]) if rec["mask"] else None,
        "pred": torch.stack(rec["pred"]) if rec["pred"] else None,
        "scores": torch.stack(rec["scores"]) if rec["scores"] else None,
        "probs": torch.stack(rec["probs"]) if rec["probs"] else None,
        # full tensors for the 100-step case; a stride-13 flat subsample of the two big weights otherwise
        "final_state": {k: (v.detach().clone() if (full_final or v.numel() < 2048) else v.detach().flatten()[::13].clone())
                        for k, v in sd.items()},
        "final_state_stride": 1 if full_final else 13,
        "dacp_ema_thr": tr.dacp_manager.ema_thresholds.clone(),
        "dacp_quality": tr.dacp_manager.class_quality_scores.clone(),
        "input_checksums": sums,
    }
    torch.save(fix, os.path.join(OUT, name + ".pt"))
    last = steps[-1]
    print(f"{name}: {len(steps)} steps, last={ {k: round(v, 5) for k, v in last.items() if isinstance(v, float)} }"
          f" masked={int(fix['mask'].sum()) if fix['mask'] is not None else 0}")


def ops_case(student):
    """Op-level vectors straight from the reference classes (utils.py / model.py)."""
    cfg, m_model, m_utils, _ = ref_harness.load_reference("iemocap")
    B, T = 6, 20
    c, n = synth.make_batches(B, T, 777, variable_len=True)
    rng = synth.make_rng_inputs(B, T, 778)
    aug = m_utils.DataAugmentation()
    xn = n["net_input"]["feats"]
    with ref_harness.supplied_rng(rng["n_weak"], rng["n_strong"], rng["feat_keep_u"], rng["t_start"]):
        weak = aug.weak_augment(xn)
        strong = aug.strong_augment(xn)
    with ref_harness.supplied_rng(rng["n_weak"][0], rng["n_strong"][0], rng["feat_keep_u"], rng["t_start"][:1]):
        aug.weak_augment(xn[0])
        strong2d = aug.strong_augment(xn[0])
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = m_model.SSRLModel(cfg)
    model.load_state_dict(full_state(student))
    model.eval()
    with torch.no_grad():
        emb_pad = model.student_encoder(c["net_input"]["feats"], c["net_input"]["padding_mask"])
        emb_nopad = model.student_encoder(c["net_input"]["feats"], None)
        logits = model.student_classifier(emb_pad)
        probs = torch.softmax(model.teacher_classifier(model.teacher_encoder(weak, n["net_input"]["padding_mask"])), 1)
    # DACP on a bigger synthetic prob table (covers empty class + single-member class)
    g = torch.Generator().manual_seed(5)
    P = torch.softmax(torch.randn(40, 4, generator=g) * 3, 1)
    P[:, 3] *= 1e-3
    P = P / P.sum(1, keepdim=True)  # class 3 never predicted
    dm = m_utils.DACPManager(4, cfg.EPOCHS, torch.device("cpu"))
    dm.class_quality_scores = torch.tensor([0.5, 0.62, 0.41, 0.5])
    anchors = torch.tensor([0.05, 0.30, 0.05, 0.05])
    dacp_out = []
    for e in (40, 41, 250):
        m, s, w = dm.calculate_mask(P, e, anchors)
        dacp_out.append({"epoch": e, "mask": m.clone(), "scores": s.clone(), "w": w.clone(), "ema_thr": dm.ema_thresholds.clone()})
    dm.update_class_quality_scores_epoch(dm.batch_scores_per_class)
    # ECDA forward + grads on embeddings
    ecda = m_utils.ECDALoss()
    g = torch.Generator().manual_seed(6)
    Bn = 48
    yc = torch.randint(0, 4, (Bn,), generator=g)
    yn = torch.randint(0, 4, (Bn,), generator=g)
    mu = torch.randn(4, 256, generator=g)
    fc = (mu[yc] + 0.7 * torch.randn(Bn, 256, generator=g)).abs().requires_grad_(True)
    fn = (mu[yn] + 0.9 * torch.randn(Bn, 256, generator=g)).abs().requires_grad_(True)
    msk = torch.rand(Bn, generator=g) > 0.3
    sc = torch.rand(Bn, generator=g)
    wce = torch.tensor([0.45, 0.55, 0.5, 0.62])
    ecda_out = {}
    for aware in (True, False):
        cfg.USE_CLASS_AWARE_MMD = aware
        fc.grad = fn.grad = None
        loss = ecda(fc, fn, yc, yn, msk, sc, wce)
        loss.backward()
        ecda_out["aware" if aware else "global"] = {"loss": loss.detach().clone(), "g_clean": fc.grad.clone(), "g_noisy": fn.grad.clone()}
    cfg.USE_CLASS_AWARE_MMD = True
    torch.save({
        "B": B, "T": T, "seed_batches": 777, "seed_rng": 778,
        "weak": weak, "strong": strong, "strong2d": strong2d,
        "emb_pad": emb_pad, "emb_nopad": emb_nopad, "logits_eval": logits, "teacher_probs": probs,
        "dacp": {"P": P, "quality0": torch.tensor([0.5, 0.62, 0.41, 0.5]), "anchors": anchors, "out": dacp_out,
                 "quality_after_epoch": dm.class_quality_scores.clone()},
        "ecda": {"yc": yc, "yn": yn, "fc": fc.detach().clone(), "fn": fn.detach().clone(), "mask": msk, "scores": sc, "wce": wce,
                 "out": ecda_out},
    }, os.path.join(OUT, "ops_small.pt"))
    print("ops_small: weak/strong/encoder/dacp/ecda vectors written")


# name -> (dataset, overrides, B, T, epochs, steps_per_epoch, seed, full_final)
CASES = {
    "step_iemocap": ("iemocap", {}, 16, 40, list(range(40, 50)), 10, 9000, True),
    "step_iemocap_warmup": ("iemocap", {}, 16, 40, [3], 5, 9500, False),
    "step_iemocap_noentropy_globalmmd": ("iemocap", {"USE_ENTROPY_IN_SCORE": False, "USE_CLASS_AWARE_MMD": False}, 16, 40, [40, 41], 5, 9600, False),
    "step_iemocap_fixedthr": ("iemocap", {"USE_DACP": False, "FIXED_CONFIDENCE_THRESHOLD": 0.6}, 16, 40, [45, 46], 5, 9700, False),
    "step_casia_default": ("casia", {}, 16, 40, [40, 41], 5, 9800, False),
    "step_casia_full": ("casia", {"USE_DACP": True, "USE_ECDA": True}, 16, 40, [40, 41], 5, 9900, False),
    "step_emodb": ("emodb", {}, 16, 40, [40, 41], 5, 10000, False),
    # BASELINE.json configs[0] / configs[1] at their exact shapes (EMODB batch 32, IEMOCAP batch 64, T=200)
    "step_emodb_b32_t200": ("emodb", {}, 32, 200, [40, 41], 3, 11000, False),
    "step_iemocap_b64_t200": ("iemocap", {}, 64, 200, [60, 61], 3, 12000, False),
}


def main(only=None):
    """No arguments: regenerate everything (pretrained student included).  With case names: ADD / refresh only those cases, starting
    from the committed tests/golden/pretrained_student.pt (so that the other fixtures, written on another host, stay untouched)."""
    os.makedirs(OUT, exist_ok=True)
    if only:
        student = torch.load(os.path.join(OUT, "pretrained_student.pt"), weights_only=False)
    else:
        student = pretrain()
        torch.save(student, os.path.join(OUT, "pretrained_student.pt"))
        ops_case(student)
    for name, (dataset, over, B, T, epochs, spe, seed, full) in CASES.items():
        if not only or name in only:
            run_case(name, dataset, student, over, B, T, epochs, spe, seed, full_final=full)


if __name__ == "__main__":
    main(sys.argv[1:])
