"""CPU restatement (PyTorch fp32 on CPU, optional fp64) of the reference's DAD training step.

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path, never the product.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import it.

Parity status: PINNED.  The reference is pure Python/PyTorch with no golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself executed in the
build container: oracle/make_golden.py imports the unmodified reference through oracle/ref_harness.py and
writes tests/golden/*.pt; tests/test_oracle_golden.py checks this file against those fixtures (and against
the live reference when /root/reference is present).  Exception: `scl_loss` has no reference
implementation (removed upstream, C/utils.py:377-378) -> "parity unpinned" for that function only.

Every function cites the reference file:line it follows.  I/ = IEMOCAP/DAD-train-IEMOCAP/.
RNG is never drawn here: noise, masks and starts are inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

PARAM_KEYS = (
    "student_encoder.pre_net.weight", "student_encoder.pre_net.bias",
    "student_classifier.fc_layer.weight", "student_classifier.fc_layer.bias",
)
TEACHER_KEYS = tuple(k.replace("student_", "teacher_") for k in PARAM_KEYS)


# ----------------------------------------------------------------------------- hyper-parameters
@dataclass
class HParams:
    """The cfg fields the step reads (I/config.py:27-148); defaults are the IEMOCAP values."""
    NUM_CLASSES: int = 4
    DROPOUT_RATE: float = 0.1
    EMA_MOMENTUM: float = 0.995
    WARMUP_EPOCHS: int = 30
    ECDA_START_EPOCH: int = 30
    SCL_START_EPOCH: int = 5001
    DACP_QUALITY_SMOOTHING_BETA: float = 0.9
    DACP_SENSITIVITY_K: float = 10.0
    DACP_QUANTILE_START: float = 0.4
    DACP_QUANTILE_END: float = 0.80
    DACP_CALIBRATION_STRENGTH_LAMBDA: float = 0.9
    DACP_THRESHOLD_SMOOTHING_ALPHA: float = 0.9
    USE_DACP: bool = True
    USE_ECDA: bool = True
    FIXED_CONFIDENCE_THRESHOLD: float = 0.9
    USE_ENTROPY_IN_SCORE: bool = True
    USE_CLASS_AWARE_MMD: bool = True
    ECDA_CLASS_ATTENTION_LAMBDA: float = 1.0
    ECDA_COMPACTNESS_WEIGHT_GAMMA: float = 0.1
    ECDA_REPULSION_WEIGHT_DELTA: float = 0.1
    WEIGHT_CONSISTENCY: float = 1.0
    TARGET_SCL_WEIGHT: float = 0.0
    WEIGHT_ECDA: float = 0.3
    EPOCHS: int = 500
    LEARNING_RATE: float = 5e-4
    WEIGHT_DECAY: float = 1e-5
    USE_LABEL_SMOOTHING: bool = True
    LABEL_SMOOTHING_FACTOR: float = 0.05
    WEAK_NOISE_STD: float = 0.01
    STRONG_NOISE_STD: float = 0.05
    TEMPORAL_MASK_RATIO: float = 0.1
    PROGRESSIVE_TRAINING: bool = True
    INITIAL_CONSISTENCY_WEIGHT: float = 0.1
    FINAL_CONSISTENCY_WEIGHT: float = 0.3
    WEIGHT_RAMP_EPOCHS: int = 30
    GRADIENT_CLIPPING: bool = True
    MAX_GRAD_NORM: float = 1.0

    @classmethod
    def from_cfg(cls, cfg):
        kw = {f: getattr(cfg, f) for f in cls.__dataclass_fields__ if hasattr(cfg, f)}
        return cls(**kw)


# ----------------------------------------------------------------------------- augmentation
def weak_augment(x, n_weak, std):
    """I/utils.py:328-331: x + randn*std, two fp32 roundings (mul, add)."""
    return x + n_weak * std


def temporal_mask_len(T, ratio):
    """I/utils.py:366: int(T_padded * ratio)."""
    return int(T * ratio)


def strong_augment(x, n_strong, std, feat_keep_u, drop_rate, t_start, mask_ratio):
    """I/utils.py:333-375.  feat_keep_u: uniform [D] draw (keep = u > rate, shared over B,T, no rescale);
    t_start[b] in [0, max(1, T-L+1)); rows [s, s+L) are assigned +0."""
    a = x + n_strong * std
    if drop_rate > 0:
        a = a * (feat_keep_u > drop_rate).to(a.dtype)
    if mask_ratio > 0 and a.dim() >= 2:
        T = a.shape[-2]
        L = temporal_mask_len(T, mask_ratio)
        if L > 0:
            a = a.clone()
            if a.dim() == 2:
                s = int(t_start[0])
                a[s:s + L] = 0
            else:
                for b in range(a.shape[0]):
                    s = int(t_start[b])
                    a[b, s:s + L] = 0
    return a


# ----------------------------------------------------------------------------- model
def encoder(x, pad, W, b):
    """I/model.py:18-41: relu(x W^T + b), zero padded frames, masked mean over T (len clamped >= 1)."""
    h = torch.relu(F.linear(x, W, b))
    if pad is None:
        return h.mean(dim=1)
    v = 1 - pad.unsqueeze(-1).to(h.dtype)
    h = h * v
    n = (1 - pad.to(h.dtype)).sum(dim=1, keepdim=True)
    return h.sum(dim=1) / torch.clamp(n, min=1.0)


def classifier(emb, keep, p, Wc, bc):
    """I/model.py:54-64: dropout (ATen: x * (bernoulli/(1-p))) then linear.  keep=None -> no dropout."""
    u = emb if (keep is None or p == 0.0) else emb * (keep / (1.0 - p))
    return F.linear(u, Wc, bc)


def ce_label_smoothing(logits, y, eps):
    """I/train.py:364,400: nn.CrossEntropyLoss(label_smoothing=eps), mean over the batch."""
    return F.cross_entropy(logits, y, label_smoothing=eps)


# ----------------------------------------------------------------------------- DACP
def certainty_scores(probs, use_entropy=True):
    """I/utils.py:400-428."""
    maxp, pred = torch.max(probs, dim=1)
    if not use_entropy:
        return maxp, pred
    ent = -(probs * torch.log2(probs + 1e-8)).sum(dim=1)
    nent = ent / math.log2(probs.size(1))
    return maxp * (1 - nent), pred


@dataclass
class DACPState:
    """I/utils.py:384-398 (+ the per-epoch host lists kept as sum/count, same mean)."""
    quality: torch.Tensor  # class_quality_scores Q [C], init 0.5
    ema_thr: torch.Tensor  # ema_thresholds theta [C], init 0.5
    epoch_scores: List[List[float]] = field(default_factory=list)

    @classmethod
    def fresh(cls, C, dtype=torch.float32):
        return cls(torch.full((C,), 0.5, dtype=dtype), torch.full((C,), 0.5, dtype=dtype), [[] for _ in range(C)])


def dacp_mask(st: DACPState, probs, epoch, anchors, hp: HParams):
    """I/utils.py:449-507.  Mutates st.ema_thr, appends to st.epoch_scores.  Returns mask(bool), scores, w."""
    scores, pred = certainty_scores(probs, hp.USE_ENTROPY_IN_SCORE)
    C = hp.NUM_CLASSES
    delta = st.quality - st.quality.mean()
    w = torch.sigmoid(hp.DACP_SENSITIVITY_K * delta)
    gamma = hp.DACP_QUANTILE_START + (hp.DACP_QUANTILE_END - hp.DACP_QUANTILE_START) * (epoch / hp.EPOCHS)
    tau = torch.zeros(C, dtype=probs.dtype)
    for c in range(C):
        sc = scores[pred == c]
        tau[c] = torch.quantile(sc, gamma) if sc.numel() > 0 else st.ema_thr[c]
    tau = tau + hp.DACP_CALIBRATION_STRENGTH_LAMBDA * (w - 0.5)
    tau = torch.max(tau, anchors.to(tau.dtype))
    a = hp.DACP_THRESHOLD_SMOOTHING_ALPHA
    st.ema_thr = a * st.ema_thr + (1 - a) * tau
    mask = scores >= st.ema_thr[pred]
    for c in range(C):
        st.epoch_scores[c].extend(scores[pred == c].detach().tolist())
    return mask, scores, w


def dacp_epoch_update(st: DACPState, hp: HParams):
    """I/utils.py:430-447: Q = beta*Q + (1-beta)*mean(scores of class c this epoch); keep Q if none."""
    import numpy as np
    cur = torch.tensor([float(np.mean(np.asarray(s, dtype=np.float32))) if len(s) > 0 else st.quality[i].item()
                        for i, s in enumerate(st.epoch_scores)], dtype=torch.float32)
    b = hp.DACP_QUALITY_SMOOTHING_BETA
    st.quality = b * st.quality + (1 - b) * cur
    st.epoch_scores = [[] for _ in st.epoch_scores]


# ----------------------------------------------------------------------------- KD
def kd_loss(student_logits, teacher_probs, mask_f):
    """I/train.py:439-447: KLDivLoss(reduction='none')(log q, p).sum(1) masked mean with +1e-8."""
    logq = F.log_softmax(student_logits, dim=1)
    kl = F.kl_div(logq, teacher_probs, reduction="none").sum(dim=1)
    return (kl * mask_f).sum() / (mask_f.sum() + 1e-8)


# ----------------------------------------------------------------------------- ECDA
def _mk_mmd_terms(S, Tt, ws, wt, kernel_mul=2.0, kernel_num=5):
    """I/utils.py:521-563 (direct [n,n,D] difference form; bandwidth detached)."""
    ns, nt = S.size(0), Tt.size(0)
    Z = torch.cat([S, Tt], dim=0)
    diff = Z.unsqueeze(0) - Z.unsqueeze(1)
    L2 = (diff * diff).sum(2)
    n = ns + nt
    bw = L2.detach().sum() / (n * n - n) if n > 1 else torch.tensor(1.0)
    bw = bw / (kernel_mul ** (kernel_num // 2))
    K = sum(torch.exp(-L2 / (bw * (kernel_mul ** i) + 1e-8)) for i in range(kernel_num))
    Wss, Wtt, Wst = torch.outer(ws, ws), torch.outer(wt, wt), torch.outer(ws, wt)
    tss = (K[:ns, :ns] * Wss).sum() / (Wss.sum() + 1e-8)
    ttt = (K[ns:, ns:] * Wtt).sum() / (Wtt.sum() + 1e-8)
    tst = (K[:ns, ns:] * Wst).sum() / (Wst.sum() + 1e-8)
    return tss, ttt, tst


def ecda_loss(clean, noisy, y_clean, y_noisy, mask, scores, w, hp: HParams):
    """I/utils.py:565-652.  `mask` bool, or float (USE_DACP=False path, re-thresholded > FIXED thr, :572-576);
    `w` has shape [C] (DACP) or [B] (fixed-threshold path, I/train.py:420)."""
    total = torch.zeros((), dtype=clean.dtype)
    ncls = w.size(0)
    if mask.dtype != torch.bool:
        mask = mask > hp.FIXED_CONFIDENCE_THRESHOLD
    if hp.USE_CLASS_AWARE_MMD:
        cents = []
        for c in range(ncls):
            f = noisy[(y_noisy == c) & mask]
            if len(f) > 0:
                cents.append(f.mean(dim=0))
        rep = torch.zeros((), dtype=clean.dtype)
        if len(cents) > 1:
            rep = -torch.pdist(torch.stack(cents), p=2).mean()
        att = torch.exp(hp.ECDA_CLASS_ATTENTION_LAMBDA * (w.mean() - w))
        for c in range(ncls):
            cs = clean[y_clean == c]
            sel = (y_noisy == c) & mask
            nz = noisy[sel]
            if len(cs) < 2 or len(nz) < 2:
                continue
            tss, ttt, tst = _mk_mmd_terms(cs, nz, torch.ones(len(cs), dtype=clean.dtype), scores[sel])
            mmd = tss + ttt - 2 * tst
            mu = nz.mean(dim=0)
            comp = ((nz - mu) ** 2).sum(dim=1).mean()
            total = total + att[c] * (mmd + hp.ECDA_COMPACTNESS_WEIGHT_GAMMA * comp + hp.ECDA_REPULSION_WEIGHT_DELTA * rep)
    else:
        nz = noisy[mask]
        if len(clean) >= 2 and len(nz) >= 2:
            tss, ttt, tst = _mk_mmd_terms(clean, nz, torch.ones(len(clean), dtype=clean.dtype),
                                          torch.ones(len(nz), dtype=clean.dtype))
            total = tss + ttt - 2 * tst
    return total


def ecda_loss_gram(clean, noisy, y_clean, y_noisy, mask, scores, w, hp: HParams):
    """Fixed-shape Gram-form restatement of ecda_loss for batches where the
    reference's [n,n,256] broadcast does not fit (global batch 4096).  Same maths as I/utils.py:521-652 with
    L2 = |zi|^2+|zj|^2-2 zi.zj; run it in fp64.  Validated against ecda_loss in tests at B <= 256."""
    if mask.dtype != torch.bool:
        mask = mask > hp.FIXED_CONFIDENCE_THRESHOLD
    dt = clean.dtype
    ncls = w.size(0)
    total = torch.zeros((), dtype=dt)

    def terms(S, Tt, ws, wt):
        Z = torch.cat([S, Tt], 0)
        ns, n = S.size(0), S.size(0) + Tt.size(0)
        sq = (Z * Z).sum(1)
        L2 = (sq[:, None] + sq[None, :] - 2 * Z @ Z.t()).clamp_min(0)
        L2 = L2 - torch.diag(torch.diag(L2))
        bw = (L2.detach().sum() / (n * n - n) if n > 1 else torch.tensor(1.0, dtype=dt)) / 4.0
        K = sum(torch.exp(-L2 / (bw * (2.0 ** i) + 1e-8)) for i in range(5))
        Wss, Wtt, Wst = torch.outer(ws, ws), torch.outer(wt, wt), torch.outer(ws, wt)
        return ((K[:ns, :ns] * Wss).sum() / (Wss.sum() + 1e-8), (K[ns:, ns:] * Wtt).sum() / (Wtt.sum() + 1e-8),
                (K[:ns, ns:] * Wst).sum() / (Wst.sum() + 1e-8))

    if hp.USE_CLASS_AWARE_MMD:
        cents = [noisy[(y_noisy == c) & mask].mean(0) for c in range(ncls) if ((y_noisy == c) & mask).any()]
        rep = -torch.pdist(torch.stack(cents), p=2).mean() if len(cents) > 1 else torch.zeros((), dtype=dt)
        att = torch.exp(hp.ECDA_CLASS_ATTENTION_LAMBDA * (w.mean() - w))
        for c in range(ncls):
            cs, sel = clean[y_clean == c], (y_noisy == c) & mask
            nz = noisy[sel]
            if len(cs) < 2 or len(nz) < 2:
                continue
            a, b, cst = terms(cs, nz, torch.ones(len(cs), dtype=dt), scores[sel].to(dt))
            comp = ((nz - nz.mean(0)) ** 2).sum(1).mean()
            total = total + att[c] * (a + b - 2 * cst + hp.ECDA_COMPACTNESS_WEIGHT_GAMMA * comp
                                      + hp.ECDA_REPULSION_WEIGHT_DELTA * rep)
    else:
        nz = noisy[mask]
        if len(clean) >= 2 and len(nz) >= 2:
            a, b, cst = terms(clean, nz, torch.ones(len(clean), dtype=dt), torch.ones(len(nz), dtype=dt))
            total = a + b - 2 * cst
    return total


def scl_loss(feats, labels, member, temperature=0.1):
    """Supervised contrastive loss (Khosla et al. 2020) -- PARITY UNPINNED: removed from the reference
    (C/utils.py:377-378; old ctor C/train_CASIA.py:396 temperature=0.1).  SURVEY.md Appendix A spec:
    L2-normalised members, positives = same class j != i, mean over anchors with >= 1 positive."""
    idx = member.nonzero().flatten()
    if idx.numel() < 2:
        return torch.zeros((), dtype=feats.dtype)
    f = F.normalize(feats[idx], dim=1)
    y = labels[idx]
    sim = f @ f.t() / temperature
    eye = torch.eye(len(idx), dtype=torch.bool)
    logden = torch.logsumexp(sim.masked_fill(eye, float("-inf")), dim=1)
    pos = (y[:, None] == y[None, :]) & ~eye
    npos = pos.sum(1)
    li = -((sim - logden[:, None]) * pos).sum(1) / npos.clamp_min(1)
    has = npos > 0
    return li[has].mean() if has.any() else torch.zeros((), dtype=feats.dtype)


# ----------------------------------------------------------------------------- schedule
def loss_weights(epoch, hp: HParams):
    """I/train.py:380-395 -> (w_kd, w_ecda)."""
    if epoch < hp.WARMUP_EPOCHS:
        return 0.0, 0.0
    if hp.PROGRESSIVE_TRAINING:
        prog = min(1.0, (epoch - hp.WARMUP_EPOCHS) / hp.WEIGHT_RAMP_EPOCHS)
        w_kd = hp.INITIAL_CONSISTENCY_WEIGHT + (hp.FINAL_CONSISTENCY_WEIGHT - hp.INITIAL_CONSISTENCY_WEIGHT) * prog
    else:
        w_kd = hp.WEIGHT_CONSISTENCY
    w_ecda = hp.WEIGHT_ECDA * min(1.0, (epoch - hp.ECDA_START_EPOCH) / hp.WEIGHT_RAMP_EPOCHS) \
        if epoch >= hp.ECDA_START_EPOCH else 0.0
    return w_kd, w_ecda


def cosine_lr(lr0, epoch, total):
    """CosineAnnealingLR(T_max=EPOCHS), stepped once per epoch (I/train.py:363,519), eta_min=0 closed form."""
    return 0.5 * lr0 * (1 + math.cos(math.pi * epoch / total))


# ----------------------------------------------------------------------------- optimiser tail
def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (I/train.py:487-488): L2 of per-tensor L2s, coef=min(1, max/(norm+1e-6))."""
    norms = torch.stack([torch.linalg.vector_norm(g, 2) for g in grads])
    total = torch.linalg.vector_norm(norms, 2)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adam_step(p, g, m, v, step, lr, wd, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor maths (I/train.py:362,489): coupled L2, bias-corrected."""
    g = g.add(p, alpha=wd)
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def ema_update(t, s, momentum):
    """I/model.py:211-223: t*m + s*(1-m) -- two muls and an add, python-double scalars rounded to fp32."""
    return t * momentum + s * (1.0 - momentum)


# ----------------------------------------------------------------------------- the whole step
class OracleStepper:
    """Full step: I/train.py:397-471 (train_step) + :484-492 (backward, clip, Adam, EMA).

    state: the 8 SSRLModel tensors (cloned).  Gradients by autograd over the restated forward, exactly as the
    reference obtains them.  dataset in {'iemocap','casia','emodb'} selects the per-trainer differences of
    SURVEY.md 5.1 (EMODB: no USE_DACP/USE_ECDA switches, E/train_emodb.py:418-437)."""

    def __init__(self, state: Dict[str, torch.Tensor], hp: HParams, anchors=None, dataset="iemocap",
                 dtype=torch.float32, lr=None):
        self.hp, self.dataset, self.dtype = hp, dataset, dtype
        self.p = {k: state[k].detach().clone().to(dtype).requires_grad_(True) for k in PARAM_KEYS}
        self.t = {k: state[k].detach().clone().to(dtype) for k in TEACHER_KEYS}
        self.m = {k: torch.zeros_like(self.p[k]) for k in PARAM_KEYS}
        self.v = {k: torch.zeros_like(self.p[k]) for k in PARAM_KEYS}
        self.nstep = 0
        self.lr = hp.LEARNING_RATE if lr is None else lr
        self.dacp = DACPState.fresh(hp.NUM_CLASSES, dtype)
        self.anchors = (anchors if anchors is not None else torch.full((hp.NUM_CLASSES,), 0.05)).to(dtype)
        self.last = {}

    def state_dict(self):
        d = {k: v.detach().clone() for k, v in self.p.items()}
        d.update({k: v.clone() for k, v in self.t.items()})
        return d

    def forward_losses(self, clean, noisy, epoch, rng):
        hp, dt = self.hp, self.dtype
        W, b, Wc, bc = (self.p[k] for k in PARAM_KEYS)
        tW, tb, tWc, tbc = (self.t[k] for k in TEACHER_KEYS)
        xc = clean["net_input"]["feats"].to(dt)
        pc = clean["net_input"]["padding_mask"]
        yc = clean["labels"]
        emb_c = encoder(xc, pc, W, b)
        logits_c = classifier(emb_c, rng["drop_clean"].to(dt), hp.DROPOUT_RATE, Wc, bc)
        ce = ce_label_smoothing(logits_c, yc, hp.LABEL_SMOOTHING_FACTOR if hp.USE_LABEL_SMOOTHING else 0.0)
        zero = torch.zeros((), dtype=dt)
        out = {"supervised_ce_loss": ce, "consistency_loss": zero, "ecda_loss": zero}
        if epoch < hp.WARMUP_EPOCHS:
            out["total_loss"] = ce
            return out
        w_kd, w_ecda = loss_weights(epoch, hp)
        xn = noisy["net_input"]["feats"].to(dt)
        pn = noisy["net_input"]["padding_mask"]
        weak = weak_augment(xn, rng["n_weak"].to(dt), hp.WEAK_NOISE_STD)
        strong = strong_augment(xn, rng["n_strong"].to(dt), hp.STRONG_NOISE_STD, rng["feat_keep_u"],
                                hp.DROPOUT_RATE, rng["t_start"], hp.TEMPORAL_MASK_RATIO)
        with torch.no_grad():
            probs = F.softmax(classifier(encoder(weak, pn, tW, tb), None, 0.0, tWc, tbc), dim=1)
        use_dacp = hp.USE_DACP or self.dataset == "emodb"
        if use_dacp:
            mask, scores, wce = dacp_mask(self.dacp, probs, epoch, self.anchors, hp)
            mask_f = mask.to(dt)
            mask_arg = mask
        else:
            scores, _ = torch.max(probs, dim=1)
            mask_f = (scores >= hp.FIXED_CONFIDENCE_THRESHOLD).to(dt)
            wce = torch.ones_like(mask_f)
            mask_arg = mask_f
        emb_s = encoder(strong, pn, W, b)
        logits_s = classifier(emb_s, rng["drop_strong"].to(dt), hp.DROPOUT_RATE, Wc, bc)
        kd, ecda = zero, zero
        if mask_f.sum().item() > 1:
            kd = kd_loss(logits_s, probs, mask_f)
            use_ecda = hp.USE_ECDA or self.dataset == "emodb"
            if use_ecda and w_ecda > 0:
                ecda = ecda_loss(emb_c, emb_s, yc, probs.argmax(dim=1), mask_arg, scores, wce, hp)
        total = ce + w_kd * kd + w_ecda * ecda
        if hp.TARGET_SCL_WEIGHT > 0 and epoch >= hp.SCL_START_EPOCH:  # optional term, absent upstream (parity unpinned)
            scl = scl_loss(torch.cat([emb_c, emb_s]), torch.cat([yc, probs.argmax(dim=1)]),
                           torch.cat([torch.ones_like(yc, dtype=torch.bool), mask_f.bool()]), temperature=0.1)
            out["scl_loss"] = scl
            total = total + hp.TARGET_SCL_WEIGHT * scl
        out.update(consistency_loss=kd, ecda_loss=ecda, total_loss=total)
        self.last = dict(probs=probs, mask=mask_f.bool(), scores=scores, pred=probs.argmax(1), emb_clean=emb_c.detach(),
                         emb_strong=emb_s.detach(), weak=weak, strong=strong, wce=wce)
        return out

    def step(self, clean, noisy, epoch, rng):
        hp = self.hp
        for p in self.p.values():
            p.grad = None
        losses = self.forward_losses(clean, noisy, epoch, rng)
        losses["total_loss"].backward()
        grads = [self.p[k].grad for k in PARAM_KEYS]
        gnorm = clip_grad_norm(grads, hp.MAX_GRAD_NORM) if hp.GRADIENT_CLIPPING else None
        self.nstep += 1
        with torch.no_grad():
            for k in PARAM_KEYS:
                adam_step(self.p[k], self.p[k].grad, self.m[k], self.v[k], self.nstep, self.lr, hp.WEIGHT_DECAY)
            if epoch >= hp.WARMUP_EPOCHS:
                for ks, kt in zip(PARAM_KEYS, TEACHER_KEYS):
                    self.t[kt] = ema_update(self.t[kt], self.p[ks].detach(), hp.EMA_MOMENTUM)
        out = {k: float(v.detach()) for k, v in losses.items()}
        out["grad_norm"] = float(gnorm) if gnorm is not None else None
        return out

    def end_epoch(self, epoch):
        """I/train.py:498-499,519: per-epoch class-quality EMA (post warm-up only); caller updates lr."""
        if epoch >= self.hp.WARMUP_EPOCHS:
            dacp_epoch_update(self.dacp, self.hp)
