"""Harness that imports the UNMODIFIED reference (read-only, /root/reference) and drives its own
train_step -> backward -> clip -> Adam -> EMA on supplied tensors.

TEST INFRASTRUCTURE ONLY.  Usable only in the build container (the reference tree does not travel to
the GPU box).  Used by oracle/make_golden.py to generate tests/golden/*.pt and by the `not gpu` tests
to pin oracle/dad_oracle.py against the live reference when the tree is present.

Reference entry points driven here:
  IEMOCAP/DAD-train-IEMOCAP/train.py:397-471  (train_step)
  IEMOCAP/DAD-train-IEMOCAP/train.py:484-492  (zero_grad/backward/clip/Adam/EMA)
  IEMOCAP/DAD-train-IEMOCAP/utils.py:317-652  (DataAugmentation, DACPManager, ECDALoss)
  IEMOCAP/DAD-train-IEMOCAP/model.py:6-265    (Encoder, Classifier, SSRLModel)
"""
import contextlib
import importlib
import io
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("DAD_REFERENCE_ROOT", "/root/reference")
_DATASETS = {
    "iemocap": ("IEMOCAP/DAD-train-IEMOCAP", "config", "train", "IEMOCAPCrossDomainTrainer"),
    "casia": ("CASIA/DAD-train-CASIA", "config_casia", "train_CASIA", "FixedCASIACrossDomainTrainer"),
    "emodb": ("EMODB/DAD-train-EMODB", "config_emodb", "train_emodb", "FixedEMODBCrossDomainTrainer"),
}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, _DATASETS["iemocap"][0]))


def load_reference(dataset="iemocap"):
    """Import (cfg, model, utils, train) modules of one dataset copy of the reference."""
    sub, cfg_name, train_name, _ = _DATASETS[dataset]
    path = os.path.join(REF_ROOT, sub)
    sys.dont_write_bytecode = True
    for stub in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if stub not in sys.modules:
            sys.modules[stub] = types.ModuleType(stub)
    # the three copies use clashing module names (model, utils, config...) -> purge before switching
    for name in ("config", "config_casia", "config_emodb", "model", "utils", "train", "train_CASIA",
                 "train_emodb", "dataload_clean", "dataload_noisy", "dataload_casia_clean",
                 "dataload_casia_noisy", "dataload_emodb_clean", "dataload_emodb_noisy"):
        sys.modules.pop(name, None)
    sys.path.insert(0, path)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            cfg = importlib.import_module(cfg_name)
            model = importlib.import_module("model")
            utils = importlib.import_module("utils")
            train = importlib.import_module(train_name)
    finally:
        sys.path.remove(path)
    return cfg, model, utils, train


class _SuppliedDropout(torch.nn.Module):
    """nn.Dropout stand-in that consumes supplied keep masks (0/1) in call order: y = x*keep/(1-p)."""

    def __init__(self, p):
        super().__init__()
        self.p = p
        self.queue = []

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        keep = self.queue.pop(0)
        return x * (keep / (1.0 - self.p))  # ATen: noise=bernoulli/(1-p); out=x*noise (Dropout.cpp _dropout_impl; same value on CUDA)


@contextlib.contextmanager
def supplied_rng(n_weak, n_strong, feat_keep_u, t_start):
    """Patch torch.randn_like / rand / randint so the reference's DataAugmentation consumes supplied draws.

    Draw order (utils.py:330,338,343,371): randn_like(weak), randn_like(strong), rand(768), B x randint.
    feat_keep_u is the uniform [768] draw (mask = u > rate)."""
    q_randn = [n_weak, n_strong]
    q_start = [int(v) for v in t_start]
    o_randn, o_rand, o_randint = torch.randn_like, torch.rand, torch.randint

    def randn_like(t, *a, **k):
        return q_randn.pop(0).to(t.dtype)

    def rand(*a, **k):
        return feat_keep_u.clone()

    def randint(lo, hi, size, *a, **k):
        v = q_start.pop(0)
        assert lo <= v < hi, (lo, v, hi)
        return torch.tensor([v])

    torch.randn_like, torch.rand, torch.randint = randn_like, rand, randint
    try:
        yield
    finally:
        torch.randn_like, torch.rand, torch.randint = o_randn, o_rand, o_randint


class RefStepper:
    """Reference trainer built with __new__ (no datasets / pretrained files), as in SURVEY.md 8c."""

    def __init__(self, dataset="iemocap", overrides=None, state=None, anchors=None, lr=None):
        self.dataset = dataset
        self.cfg, self.m_model, self.m_utils, self.m_train = load_reference(dataset)
        cfg = self.cfg
        for k, v in (overrides or {}).items():
            setattr(cfg, k, v)
        cls = getattr(self.m_train, _DATASETS[dataset][3])
        tr = cls.__new__(cls)
        tr.device = torch.device("cpu")
        with contextlib.redirect_stdout(io.StringIO()):
            tr.model = self.m_model.SSRLModel(cfg)
        if state is not None:
            tr.model.load_state_dict(state)
        # supplied-mask dropout for the student classifier (teacher p=0 draws nothing, model.py:118-122)
        self.drop = _SuppliedDropout(cfg.DROPOUT_RATE)
        tr.model.student_classifier.dropout = self.drop
        tr.num_classes = cfg.NUM_CLASSES
        tr.WARMUP_EPOCHS = cfg.WARMUP_EPOCHS
        tr.initial_consistency_weight = cfg.INITIAL_CONSISTENCY_WEIGHT if cfg.PROGRESSIVE_TRAINING else cfg.WEIGHT_CONSISTENCY
        tr.final_consistency_weight = cfg.FINAL_CONSISTENCY_WEIGHT if cfg.PROGRESSIVE_TRAINING else cfg.WEIGHT_CONSISTENCY
        tr.current_consistency_weight = tr.initial_consistency_weight
        tr.weight_ecda = 0.0
        tr.weight_scl = 0.0
        tr.tracked_sample_indices = None
        tr.bias_analysis_log = []
        tr.calibrated_anchors = anchors if anchors is not None else torch.full((cfg.NUM_CLASSES,), 0.05)
        with contextlib.redirect_stdout(io.StringIO()):
            tr._setup_training_components()
            tr._setup_augmentation()
        if lr is not None:
            for g in tr.optimizer.param_groups:
                g["lr"] = lr
        tr.model.train()
        self.tr = tr

    def step(self, clean, noisy, epoch, rng):
        """One full optimisation step.  rng: dict(drop_clean, n_weak, n_strong, feat_keep_u, t_start, drop_strong)."""
        tr, cfg = self.tr, self.cfg
        tr.update_loss_weights(epoch)
        warm = tr.is_warmup_phase(epoch)
        self.drop.queue = [rng["drop_clean"]] + ([] if warm else [rng["drop_strong"]])
        tr.optimizer.zero_grad()
        with supplied_rng(rng["n_weak"], rng["n_strong"], rng["feat_keep_u"], rng["t_start"]):
            losses = tr.train_step(clean, noisy, epoch)
        losses["total_loss"].backward()
        gnorm = None
        if cfg.GRADIENT_CLIPPING:
            gnorm = torch.nn.utils.clip_grad_norm_(tr.model.parameters(), cfg.MAX_GRAD_NORM)
        tr.optimizer.step()
        if not warm:
            tr.model.update_teacher_ema()
        out = {k: float(v.detach()) for k, v in losses.items()}
        out["grad_norm"] = float(gnorm) if gnorm is not None else None
        return out
