"""Stage-1 supervised pre-training on clean features -- the reference's L1 layer
(IEMOCAP/pretrain-and-processed-IEMOCAP/model.py:4-21 `BaseModel`, train_for_clean.py:154-155,393-420) on the SAME kernels:

    BaseModel = Linear(768,256) -> ReLU -> padding-masked mean -> Linear(256,4), plain CrossEntropyLoss, Adam(lr 2e-4, wd 1e-5)

is exactly the DAD step's warm-up form (CE on the clean stream only) with dropout, label smoothing and gradient clipping
switched off, so one `CleanPretrainer.train_epoch` = the fused engine step per batch.  The checkpoint it writes is the file
stage 2 loads: a bare state_dict with `pre_net.{weight,bias}`, `post_net.{weight,bias}` (I/model.py:159-181)."""
import types

import torch

from . import _lib
from .engine import DADStepEngine

STAGE1_DEFAULTS = dict(LEARNING_RATE=2e-4, WEIGHT_DECAY=1e-5, MAX_EPOCHS=100, BATCH_SIZE=64)  # .../config.py:15-18


def stage1_cfg(base_cfg, **overrides):
    """A config namespace that makes the DAD engine compute the stage-1 step: never leaves warm-up, no dropout / smoothing / clip."""
    ns = types.SimpleNamespace(**{k: v for k, v in vars(base_cfg).items() if k.isupper()})
    ns.WARMUP_EPOCHS = 10 ** 9
    ns.DROPOUT_RATE = 0.0
    ns.USE_LABEL_SMOOTHING = False
    ns.GRADIENT_CLIPPING = False
    ns.LEARNING_RATE_SCHEDULER = "none"
    for k, v in {**STAGE1_DEFAULTS, **overrides}.items():
        setattr(ns, k, v)
    ns.EPOCHS = ns.MAX_EPOCHS
    return ns


class CleanPretrainer:
    """train_for_clean.py's loop: `train_one_epoch` (:393-420), validation accuracy, best-state checkpoint (:240-259)."""

    def __init__(self, base_cfg, train_loader, val_loader, capacity_T, state=None, seed=42, **overrides):
        self.cfg = stage1_cfg(base_cfg, **overrides)
        self.train_loader, self.val_loader = train_loader, val_loader
        mode = _lib.GEMM_TC if str(getattr(self.cfg, "DAD_GEMM_MODE", "tc")).lower() in ("tc", "f16x3") else _lib.GEMM_FP32
        self.engine = DADStepEngine(self.cfg, self.cfg.BATCH_SIZE, max(32, int(capacity_T)), gemm_mode=mode, noise_mode="philox", seed=seed,
                                    operand_cache="none")
        if state is None:
            from . import synth
            state = synth.init_state(seed)
        self.engine.load_state_dict(state)
        self.engine.set_lr(self.cfg.LEARNING_RATE)
        self.best_acc, self.best_state = -1.0, None

    def train_epoch(self):
        eng = self.engine
        eng.loss_acc.zero_()
        for batch in self.train_loader:
            eng.set_inputs(batch, None)
            eng.step(0)
        return eng.pop_epoch_losses()["supervised_ce_loss"]

    @torch.no_grad()
    def accuracy(self, loader):
        from . import ops
        sd = self.engine.state_dict()
        W, b = sd["student_encoder.pre_net.weight"], sd["student_encoder.pre_net.bias"]
        Wc, bc = sd["student_classifier.fc_layer.weight"], sd["student_classifier.fc_layer.bias"]
        hit = n = 0
        for batch in loader:
            x = batch["net_input"]["feats"].to(W.device)
            pad = batch["net_input"]["padding_mask"].to(W.device)
            emb, _, _ = ops.encoder_pool_fwd(x, pad, W, b, gemm_mode=self.engine._mode(x.shape[1]), want_bits=False)
            pred = (emb @ Wc.t() + bc).argmax(dim=1)
            hit += int((pred == batch["labels"].to(W.device)).sum())
            n += pred.numel()
        return hit / max(n, 1)

    def state_dict(self):
        """Bare stage-1 state_dict (pre_net.* / post_net.*) -- what SSRLModel.load_complete_pretrained_weights expects."""
        sd = self.engine.state_dict()
        return {"pre_net.weight": sd["student_encoder.pre_net.weight"].cpu(), "pre_net.bias": sd["student_encoder.pre_net.bias"].cpu(),
                "post_net.weight": sd["student_classifier.fc_layer.weight"].cpu(), "post_net.bias": sd["student_classifier.fc_layer.bias"].cpu()}

    def train(self, epochs=None, ckpt_path=None):
        hist = []
        for _ in range(epochs or self.cfg.MAX_EPOCHS):
            loss = self.train_epoch()
            acc = self.accuracy(self.val_loader)
            hist.append({"train_loss": loss, "val_acc": acc})
            if acc > self.best_acc:
                self.best_acc, self.best_state = acc, self.state_dict()
                if ckpt_path:
                    torch.save(self.best_state, ckpt_path)
        return hist
