"""Drop-in op-level surface of the reference's three hot classes (I/utils.py:317-652; identical maths in C/ and E/):

    DataAugmentation(weak_noise_std=None, strong_noise_std=None, dropout_rate=None, temporal_mask_ratio=None)
        .weak_augment(data) / .strong_augment(data)                                     I/utils.py:317-375
    DACPManager(num_classes, total_epochs, device)
        .calculate_certainty_scores(probs) [static] / .calculate_mask(teacher_probs, epoch, calibrated_anchors)
        .update_class_quality_scores_epoch(batch_scores_per_class)                        I/utils.py:379-507
    ECDALoss(kernel_type='rbf', kernel_mul=2.0, kernel_num=5)
        .forward(clean_feats, noisy_feats, clean_labels, noisy_labels, noisy_mask, noisy_scores, class_weights_wce)
                                                                                          I/utils.py:510-652
Same constructor arguments, method names, argument meaning, return types and state attributes
(`class_quality_scores`, `ema_thresholds`, `batch_scores_per_class`), every one running on the hand-written sm_100a
kernels through the C ABI.  cfg fields are read LIVE at call time (the runners override them with setattr,
I/run_granular_ablations.py:25-44; the reference re-imports config inside the methods for the same reason,
I/utils.py:410,567).  The training hot loop does not go through these classes -- it uses engine.DADStepEngine, which
fuses them with the encoder / heads / optimiser -- they serve validation, anchor calibration, analysis scripts and any
caller written against the reference API.  CUDA tensors only: there is no CPU path."""
import ctypes as C
import sys

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import D_HID, N_CLS, DacpArgs, DadError, EcdaArgs, Pack, check, lib, ptr, stream_ptr


def _cfg():
    """The config module the caller is using: the reference binds `import config as cfg` (I/utils.py:11); honour a
    top-level `config` / `config_casia` / `config_emodb` module if the application imported one, else the package's."""
    for name in ("config", "config_casia", "config_emodb"):
        m = sys.modules.get(name)
        if m is not None and hasattr(m, "WEAK_NOISE_STD"):
            return m
    from . import config
    return config


def use_config(module):
    """Pin the config module the op-level classes read (the trainers call this with their dataset's config)."""
    global _cfg
    _cfg = lambda: module  # noqa: E731


def _f32(x):
    return float(np.float32(x))


# ------------------------------------------------------------------------------------------------------ augmentation
class DataAugmentation:
    """I/utils.py:317-375.  weak: x + N(0, std_w^2); strong: x + N(0, std_s^2), one [768] feature keep-mask shared by the whole
    batch (multiply by 0/1, no rescale), then a per-sample temporal span of int(T_padded * ratio) frames set to +0.
    Noise comes from the in-kernel Philox field keyed by a seed drawn from torch's default generator (so torch.manual_seed
    makes runs reproducible); the feature mask and the span starts are drawn with the reference's own torch calls."""

    def __init__(self, weak_noise_std=None, strong_noise_std=None, dropout_rate=None, temporal_mask_ratio=None):
        cfg = _cfg()
        # `is not None`, as the reference (I/utils.py:322-326): an explicit 0 / 0.0 disables that augmentation (ablations)
        self.weak_noise_std = weak_noise_std if weak_noise_std is not None else cfg.WEAK_NOISE_STD
        self.strong_noise_std = strong_noise_std if strong_noise_std is not None else cfg.STRONG_NOISE_STD
        self.dropout_rate = dropout_rate if dropout_rate is not None else cfg.DROPOUT_RATE
        self.temporal_mask_ratio = temporal_mask_ratio if temporal_mask_ratio is not None else cfg.TEMPORAL_MASK_RATIO

    @staticmethod
    def _seed():
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def weak_augment(self, data):
        return ops.augment(data.float(), _f32(self.weak_noise_std), seed=self._seed(), stream=_lib.STREAM_WEAK)

    def strong_augment(self, data):
        x = data.float()
        T = x.shape[-2] if x.dim() >= 2 else 1
        B = x.shape[0] if x.dim() == 3 else 1
        keep = None
        if self.dropout_rate > 0:
            keep = (torch.rand(x.size(-1), device=x.device) > self.dropout_rate).float()          # I/utils.py:343
        L = int(T * self.temporal_mask_ratio)                                                     # I/utils.py:366
        t_start = None
        if self.temporal_mask_ratio > 0 and L > 0:
            t_start = torch.randint(0, max(1, T - L + 1), (B,)).to(torch.int32).to(x.device)      # CPU generator, I/utils.py:371
        return ops.augment(x, _f32(self.strong_noise_std), seed=self._seed(), stream=_lib.STREAM_STRONG, feat_keep=keep, t_start=t_start,
                           mask_len=L, strong=True)


# ------------------------------------------------------------------------------------------------------ DACP
class DACPManager:
    """I/utils.py:379-507.  State lives on the device: `class_quality_scores` Q [C] (init 0.5), `ema_thresholds` theta [C]
    (init 0.5); `batch_scores_per_class` keeps the reference's per-class host lists (filled lazily from device buffers)."""

    def __init__(self, num_classes, total_epochs, device):
        if num_classes != N_CLS:
            raise DadError(f"kernels are specialised for {N_CLS} classes (cfg.NUM_CLASSES), got {num_classes}")
        self.num_classes, self.total_epochs, self.device = num_classes, total_epochs, torch.device(device)
        self.class_quality_scores = torch.full((num_classes,), 0.5, device=self.device)
        self.ema_thresholds = torch.full((num_classes,), 0.5, device=self.device)
        self.batch_scores_per_class = [[] for _ in range(num_classes)]
        self._epoch_sum = torch.zeros(num_classes, dtype=torch.float64, device=self.device)
        self._epoch_cnt = torch.zeros(num_classes, dtype=torch.int64, device=self.device)

    @staticmethod
    def calculate_certainty_scores(probs):
        """-> (certainty_scores [B] f32, pseudo_labels [B] i64).  I/utils.py:400-428 (cfg.USE_ENTROPY_IN_SCORE read live)."""
        _lib.require_device()
        p = probs.detach().float().contiguous()
        if p.dim() != 2 or p.shape[1] != N_CLS:
            raise DadError(f"probs must be [B,{N_CLS}], got {tuple(probs.shape)}")
        B = p.shape[0]
        scores = torch.empty(B, dtype=torch.float32, device=p.device)
        preds = torch.empty(B, dtype=torch.int32, device=p.device)
        use_entropy = int(bool(getattr(_cfg(), "USE_ENTROPY_IN_SCORE", True)))
        check(lib().dad_certainty_scores(ptr(p, torch.float32, "probs"), B, use_entropy, ptr(scores), ptr(preds), stream_ptr()))
        return scores, preds.long()

    def update_class_quality_scores_epoch(self, batch_scores_per_class):
        """Q = beta*Q + (1-beta)*mean(scores of class c this epoch); Q kept where a class saw no sample (I/utils.py:430-447)."""
        beta = _cfg().DACP_QUALITY_SMOOTHING_BETA
        q = self.class_quality_scores.tolist()
        cur = [float(np.mean(np.asarray(s, dtype=np.float32))) if len(s) > 0 else q[i] for i, s in enumerate(batch_scores_per_class)]
        cur = torch.tensor(cur, dtype=torch.float32, device=self.device)
        self.class_quality_scores = beta * self.class_quality_scores + (1 - beta) * cur
        self.batch_scores_per_class = [[] for _ in range(self.num_classes)]
        self._epoch_sum.zero_(); self._epoch_cnt.zero_()

    def calculate_mask(self, teacher_probs, epoch, calibrated_anchors):
        """-> (mask [B] bool, certainty_scores [B], class_weights_wce [C]).  Advances `ema_thresholds` (I/utils.py:449-507)."""
        cfg = _cfg()
        scores, preds = self.calculate_certainty_scores(teacher_probs)
        B, dev = scores.shape[0], scores.device
        pack = torch.zeros(_lib.pack_rank_words(B), dtype=torch.float32, device=dev)
        gv = _lib.pack_global_views(pack, B, 1)
        gv["score"][0].copy_(scores)
        gv["pred"][0].copy_(preds.int())
        mask = torch.zeros(B, dtype=torch.uint8, device=dev)
        out_w, out_thr = torch.empty(N_CLS, device=dev), torch.empty(N_CLS, device=dev)
        nmask = torch.zeros(1, dtype=torch.int32, device=dev)
        anchors = (calibrated_anchors if calibrated_anchors is not None else torch.zeros(N_CLS)).to(dev, torch.float32).contiguous()
        a = DacpArgs()
        a.pack, a.use_dacp = Pack(pack.data_ptr(), B, 1), 1
        a.gamma = _f32(cfg.DACP_QUANTILE_START + (cfg.DACP_QUANTILE_END - cfg.DACP_QUANTILE_START) * (epoch / self.total_epochs))
        a.sens_k, a.calib_lambda = _f32(cfg.DACP_SENSITIVITY_K), _f32(cfg.DACP_CALIBRATION_STRENGTH_LAMBDA)
        a.alpha, a.one_minus_alpha = _f32(cfg.DACP_THRESHOLD_SMOOTHING_ALPHA), _f32(1 - cfg.DACP_THRESHOLD_SMOOTHING_ALPHA)
        a.fixed_thr = _f32(getattr(cfg, "FIXED_CONFIDENCE_THRESHOLD", 0.9))
        self.class_quality_scores = self.class_quality_scores.float().contiguous()
        self.ema_thresholds = self.ema_thresholds.float().contiguous()
        a.quality, a.anchors, a.ema_thr = self.class_quality_scores.data_ptr(), anchors.data_ptr(), self.ema_thresholds.data_ptr()
        a.mask, a.out_w, a.out_thr, a.out_nmask = mask.data_ptr(), out_w.data_ptr(), out_thr.data_ptr(), nmask.data_ptr()
        a.epoch_sum, a.epoch_cnt = self._epoch_sum.data_ptr(), self._epoch_cnt.data_ptr()
        check(lib().dad_dacp_mask(C.byref(a), stream_ptr()))
        # the reference appends each class's scores to host lists every step (I/utils.py:503-505, one sync per class);
        # keep that contract for callers that read the lists
        s_host, p_host = scores.tolist(), preds.tolist()
        for s, c in zip(s_host, p_host):
            self.batch_scores_per_class[c].append(s)
        return mask.bool(), scores, out_w


# ------------------------------------------------------------------------------------------------------ ECDA
class _EcdaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, clean, noisy, y_clean, y_noisy, mask_u8, scores, w_class, class_aware, lambda_cls, gamma_comp, delta_rep):
        # the reference accepts clean / noisy batches of different size (I/utils.py:565-652 indexes them separately): the shorter
        # one is padded with null samples (label / pseudo-label -1, mask 0) that belong to no class and receive no gradient
        Bc, Bn, dev = clean.shape[0], noisy.shape[0], clean.device
        B = max(Bc, Bn, 1)
        pack = torch.zeros(_lib.pack_rank_words(B), dtype=torch.float32, device=dev)
        gv = _lib.pack_global_views(pack, B, 1)
        gv["pred"][0].fill_(-1)
        gv["label"][0].fill_(-1)
        gv["emb_clean"][0][:Bc].copy_(clean)
        gv["emb_strong"][0][:Bn].copy_(noisy)
        gv["score"][0][:Bn].copy_(scores)
        gv["pred"][0][:Bn].copy_(y_noisy.int())
        gv["label"][0][:Bc].copy_(y_clean.int())
        if Bn < B:
            mask_u8 = torch.cat([mask_u8, torch.zeros(B - Bn, dtype=torch.uint8, device=dev)])
        nmask = mask_u8.sum().to(torch.int32).reshape(1)
        g_clean, g_strong = torch.zeros(B, D_HID, device=dev), torch.zeros(B, D_HID, device=dev)
        loss3 = torch.zeros(4, device=dev)
        ws = torch.zeros(lib().dad_ecda_workspace_bytes(B, 1), dtype=torch.uint8, device=dev)
        a = EcdaArgs()
        a.pack, a.rank, a.class_aware = Pack(pack.data_ptr(), B, 1), 0, int(class_aware)
        a.mask, a.nmask, a.w_class = mask_u8.data_ptr(), nmask.data_ptr(), w_class.data_ptr()
        a.lambda_cls, a.gamma_comp, a.delta_rep, a.w_ecda = lambda_cls, gamma_comp, delta_rep, 1.0
        a.g_clean, a.g_strong, a.loss_out = g_clean.data_ptr(), g_strong.data_ptr(), loss3.data_ptr()
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        check(lib().dad_ecda(C.byref(a), stream_ptr()))
        ctx.save_for_backward(g_clean[:Bc], g_strong[:Bn])
        return loss3[2].clone()

    @staticmethod
    def backward(ctx, g):
        g_clean, g_strong = ctx.saved_tensors
        return (g * g_clean, g * g_strong) + (None,) * 9


class ECDALoss(nn.Module):
    """I/utils.py:510-652: class-aware, attention-weighted 5-kernel Gaussian MMD + compactness + centroid repulsion
    (cfg.USE_CLASS_AWARE_MMD=False: one global MMD, I/utils.py:633-650), forward and analytic backward in two kernels
    (csrc/k_ecda.cu).  The kernels skip everything when fewer than 2 noisy samples are masked in; the reference's forward
    returns 0 in that case as well (every class needs >= 2 masked noisy members, I/utils.py:609, and so does the global branch,
    :641), so the gate changes no value.  Clean and noisy batches may differ in size, as upstream."""

    def __init__(self, kernel_type="rbf", kernel_mul=2.0, kernel_num=5):
        super().__init__()
        if (kernel_type, float(kernel_mul), int(kernel_num)) != ("rbf", 2.0, 5):
            raise DadError("ECDALoss kernels are specialised for the reference's rbf / kernel_mul=2.0 / kernel_num=5")
        self.kernel_type, self.kernel_mul, self.kernel_num = kernel_type, kernel_mul, kernel_num

    def forward(self, clean_feats, noisy_feats, clean_labels, noisy_labels, noisy_mask, noisy_scores, class_weights_wce):
        _lib.require_device()
        cfg = _cfg()
        if clean_feats.dim() != 2 or noisy_feats.dim() != 2 or clean_feats.shape[1] != D_HID or noisy_feats.shape[1] != D_HID:
            raise DadError(f"ECDALoss: clean/noisy feats must be [B,{D_HID}] (got {tuple(clean_feats.shape)}, {tuple(noisy_feats.shape)})")
        mask = noisy_mask
        if mask.dtype != torch.bool:  # float mask of the USE_DACP=False path, re-thresholded as in I/utils.py:572-576
            mask = mask > getattr(cfg, "FIXED_CONFIDENCE_THRESHOLD", 0.9)
        w = class_weights_wce.detach().float()
        if w.numel() != N_CLS:  # fixed-threshold path passes ones of shape [B] (I/train.py:420): class attention is exp(0) = 1
            w = torch.ones(N_CLS, device=clean_feats.device)
        return _EcdaFn.apply(clean_feats.float().contiguous(), noisy_feats.float().contiguous(), clean_labels, noisy_labels,
                             mask.to(torch.uint8).contiguous(), noisy_scores.detach().float().contiguous(), w.contiguous(),
                             bool(getattr(cfg, "USE_CLASS_AWARE_MMD", True)), _f32(cfg.ECDA_CLASS_ATTENTION_LAMBDA),
                             _f32(cfg.ECDA_COMPACTNESS_WEIGHT_GAMMA), _f32(cfg.ECDA_REPULSION_WEIGHT_DELTA))
