"""Cross-domain DAD trainer: the reference's L4 layer (I/train.py:44-803, C/train_CASIA.py, E/train_emodb.py) on the fused engine.

Same public surface as the reference trainers:
    Trainer(fold=, experiment_name=None).train() -> {'best_noisy_weighted_acc', 'results_dir'}
    .train_step(clean_batch, noisy_batch, epoch) -> {'total_loss','supervised_ce_loss','consistency_loss','ecda_loss'[,'scl_loss']}
    .train_epoch(epoch), .validate(loader, domain_name), .update_loss_weights(epoch), .is_warmup_phase(epoch)
    attributes: model, optimizer, scheduler, dacp_manager, ecda_criterion, augmenter, training_history, calibrated_anchors, ...
    files: <results>/reports/BEST_detailed_results_epoch_*.json (summary.noisy.{w_acc,w_f1} as "63.21%", parsed by the ablation
    runners, I/run_granular_ablations.py:62-67), models/<ds>_cross_domain_best.pth, training_history.json, FINAL_test_set_results.json

What differs, deliberately.  In the reference `train_step` only builds the loss and `train_epoch` calls backward / clip / Adam /
EMA (I/train.py:484-492).  Here ONE call of `train_step` enqueues the whole iteration -- forward, losses, analytic backward, clip,
Adam, EMA -- as a CUDA graph replay of hand-written kernels (engine.DADStepEngine); the returned loss tensors are detached device
scalars (no `.backward()` to call), and `train_epoch` performs no per-step host synchronisation: epoch means come from device
accumulators (the reference does 5 `.item()` per step, I/train.py:494-495)."""
import json
import logging
import math
import os
import re
from collections import defaultdict
from datetime import datetime

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib, utils
from .engine import DADStepEngine
from .model import SSRLModel
from .utils import DACPManager, DataAugmentation, ECDALoss

logger = logging.getLogger("dad_b200.train")
if not logging.getLogger().handlers:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")


class EngineOptimizer:
    """`trainer.optimizer` facade over the engine's fused Adam (torch.optim.Adam(lr, weight_decay), I/train.py:362): exposes
    param_groups / state_dict() / load_state_dict() in torch's layout (per-parameter exp_avg, exp_avg_sq, step), built over all 8
    model tensors like the reference (teacher tensors carry no state)."""

    def __init__(self, engine, cfg):
        self.engine = engine
        self.param_groups = [{"lr": cfg.LEARNING_RATE, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": cfg.WEIGHT_DECAY,
                              "amsgrad": False, "maximize": False, "params": list(range(8))}]

    def zero_grad(self, set_to_none=True):  # gradients never persist between fused steps
        pass

    def state_dict(self):
        e = self.engine
        step = float(e.step_count.item())
        state = {}
        for i, (m, v) in enumerate(zip(e.views(e.m), e.views(e.v))):
            state[i] = {"step": torch.tensor(step), "exp_avg": m.detach().clone(), "exp_avg_sq": v.detach().clone()}
        return {"state": state, "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        e = self.engine
        for i, (m, v) in enumerate(zip(e.views(e.m), e.views(e.v))):
            if i in sd["state"]:
                m.copy_(sd["state"][i]["exp_avg"]); v.copy_(sd["state"][i]["exp_avg_sq"])
                e.step_count.fill_(int(float(sd["state"][i]["step"])))
        self.param_groups[0].update({k: v for k, v in sd["param_groups"][0].items() if k != "params"})


class CosineScheduler:
    """CosineAnnealingLR(T_max=EPOCHS, eta_min=0) stepped once per epoch (I/train.py:363,519), closed form."""

    def __init__(self, optimizer, lr0, t_max):
        self.optimizer, self.lr0, self.t_max, self.last_epoch = optimizer, lr0, t_max, 0

    def lr_at(self, epoch):
        return 0.5 * self.lr0 * (1 + math.cos(math.pi * epoch / self.t_max))

    def step(self):
        self.last_epoch += 1
        self.optimizer.param_groups[0]["lr"] = self.lr_at(self.last_epoch)

    def get_last_lr(self):
        return [self.optimizer.param_groups[0]["lr"]]


class CrossDomainTrainer:
    """Dataset-independent body; subclasses set DATASET, cfg module, results dir names and the loader factory."""

    DATASET = "iemocap"          # engine dataset switch (SURVEY.md 5.1 per-trainer differences)
    DISPLAY = "IEMOCAP"
    RESULTS_BASE = "iemocap_mutil-noisy_cross_domain_results"
    BEST_NAME = "iemocap_cross_domain_best.pth"
    WEIGHT_FIELD = "PRETRAINED_IEMOCAP_WEIGHT"
    N_FOLDS_TOTAL = 5
    EXTRA_LOSS_KEYS = ()         # CASIA / EMODB return an always-zero 'scl_loss' (C/train_CASIA.py:442,503-504)
    REAL_NOISE_DIRS = False      # CASIA / EMODB also parse "processed_features_<noise>_<n>db" noisy dirs
    TRACK_BIAS = False           # IEMOCAP only: confirmation-bias log + DACP/ECDA history (I/train.py:424-437,501-517)
    cfg = None

    # ------------------------------------------------------------------ construction
    def __init__(self, fold=0, experiment_name=None, loaders=None, require_pretrained=True):
        cfg = self.cfg
        self.experiment_name = experiment_name
        self.fold = fold
        if not (torch.cuda.is_available() and getattr(cfg, "USE_CUDA", True)):
            raise _lib.DadError("dad_b200 trainers need a CUDA device (B200, sm_100a): there is no CPU training path")
        self.device = torch.device("cuda", torch.cuda.current_device())
        utils.use_config(cfg)
        self.noise_info = self._extract_noise_info()
        self.results_dir = self._setup_results_directory(experiment_name)
        self.WARMUP_EPOCHS = cfg.WARMUP_EPOCHS
        self.target_ecda_weight = cfg.WEIGHT_ECDA
        self.weight_ecda = 0.0
        self.initial_consistency_weight = cfg.INITIAL_CONSISTENCY_WEIGHT if cfg.PROGRESSIVE_TRAINING else cfg.WEIGHT_CONSISTENCY
        self.final_consistency_weight = cfg.FINAL_CONSISTENCY_WEIGHT if cfg.PROGRESSIVE_TRAINING else cfg.WEIGHT_CONSISTENCY
        self.current_consistency_weight = self.initial_consistency_weight
        self.best_clean_acc = self.best_noisy_acc = 0.0
        self.best_clean_weighted_f1 = self.best_noisy_weighted_f1 = 0.0
        self.best_clean_weighted_acc = self.best_noisy_weighted_acc = 0.0
        self.best_results = {"epoch": 0, "clean_results": None, "noisy_results": None, "clean_confusion_matrix": None,
                             "noisy_confusion_matrix": None}
        self.training_history = defaultdict(list)
        self.patience_counter = 0
        self.tracked_sample_indices = None
        self.bias_analysis_log = []
        self._bias_pending = []
        self.dacp_manager = None
        self.calibrated_anchors = None
        if loaders is None and getattr(cfg, "DAD_SYNTHETIC_DATA", False):
            # offline stand-in data, selectable the way the runners select everything else: setattr on the config module
            # (I/run_granular_ablations.py:25-30) -- no dataset files, no stage-1 checkpoint
            loaders = self.synthetic_loaders(fold, device=self._data_device_for(cfg))
            require_pretrained = False
        self._loaders = loaders
        self._require_pretrained = require_pretrained
        logger.info(f"{self.DISPLAY} cross-domain trainer (DACP+ECDA), fold {fold + 1}/{self.N_FOLDS_TOTAL}, noise {self.noise_info['display_name']}, "
                    f"results -> {self.results_dir}")
        self._setup_components()

    def _extract_noise_info(self):
        """root type / noise type / dB from cfg.NOISY_DATA_DIR (I/train.py:113-192)."""
        path = self.cfg.NOISY_DATA_DIR
        if self.REAL_NOISE_DIRS:  # CASIA / EMODB: processed_features_<noise>_<n>db (C/train_CASIA.py:122-142)
            m = re.search(r"processed_features_([^_]+)_(\d+)db", path, re.IGNORECASE)
            if m:
                return {"root_type": "real_noise", "noise_type": m.group(1), "db_value": f"{m.group(2)}db", "display_name": f"real-{m.group(1)}-{m.group(2)}db"}
        for pat in (r"root1-([^.]+)\.wav-(\d+)db", r"root1-([^-]+)-(\d+)db"):
            m = re.search(pat, path, re.IGNORECASE)
            if m:
                return {"root_type": "root1", "noise_type": m.group(1), "db_value": f"{m.group(2)}db", "display_name": f"root1-{m.group(1)}-{m.group(2)}db"}
        m = re.search(r"root2-(\d+)db", path, re.IGNORECASE)
        if m:
            return {"root_type": "root2", "noise_type": None, "db_value": f"{m.group(1)}db", "display_name": f"root2-{m.group(1)}db"}
        for pat in (r"(\d+)db", r"(-?\d+)_?db"):
            m = re.search(pat, path, re.IGNORECASE)
            if m:
                return {"root_type": "unknown", "noise_type": "unknown", "db_value": f"{m.group(1)}db", "display_name": f"unknown-{m.group(1)}db"}
        return {"root_type": "unknown", "noise_type": "unknown", "db_value": "unknown_db", "display_name": "unknown-unknown-unknown_db"}

    def _setup_results_directory(self, experiment_name=None):
        """base[/experiment]/root1/<noise>/<db>/fold_k | base/root2/<db>/fold_k | base/unknown/fold_k (I/train.py:194-245)."""
        base = self.RESULTS_BASE
        if experiment_name:
            base = os.path.join(base, re.sub(r'[\\/*?:"<>|]', "", experiment_name))
        ni = self.noise_info
        if ni["root_type"] in ("root1", "real_noise"):
            d = os.path.join(base, ni["root_type"], ni["noise_type"], ni["db_value"], f"fold_{self.fold + 1}")
        elif ni["root_type"] == "root2":
            d = os.path.join(base, "root2", ni["db_value"], f"fold_{self.fold + 1}")
        else:
            d = os.path.join(base, "unknown", f"fold_{self.fold + 1}")
        for sub in ("", "models", "plots", "reports"):
            os.makedirs(os.path.join(d, sub), exist_ok=True)
        return d

    def _setup_components(self):
        self._setup_dataloaders()
        self._setup_model()
        self._load_pretrained_weights()
        if getattr(self.cfg, "ANCHOR_CALIBRATION_ENABLED", False):
            self._run_anchor_calibration()
        self._setup_training_components()
        self._setup_augmentation()

    def _make_loaders(self, batch_size):
        """-> ((clean_train, clean_val, clean_test, class_names, num_classes), (student, teacher, noisy_val, noisy_test))"""
        raise NotImplementedError

    @staticmethod
    def _data_device_for(cfg):
        return "cuda" if getattr(cfg, "DAD_DEVICE_DATA", True) else None

    def _data_device(self):
        """Device the loaders keep the splits on (None = pinned host memory + one H2D copy per batch)."""
        return self.device if getattr(self.cfg, "DAD_DEVICE_DATA", True) else None

    def _setup_dataloaders(self):
        clean, noisy = self._loaders(self.cfg.BATCH_SIZE) if self._loaders is not None else self._make_loaders(self.cfg.BATCH_SIZE)
        self.clean_train_loader, self.clean_val_loader, self.clean_test_loader, self.class_names, self.num_classes = clean
        self.noisy_student_loader, self.noisy_teacher_loader, self.noisy_val_loader, self.noisy_test_loader = noisy
        for ld in (self.clean_train_loader, self.noisy_student_loader, self.noisy_teacher_loader):
            if hasattr(ld, "t_bucket"):
                ld.t_bucket = int(getattr(self.cfg, "DAD_T_BUCKET", 0))  # 0: the reference's per-batch padding (dataload.UtteranceLoader.t_bucket)
        logger.info(f"loaders: clean {len(self.clean_train_loader)}/{len(self.clean_val_loader)}/{len(self.clean_test_loader)} batches, "
                    f"noisy {len(self.noisy_student_loader)}/{len(self.noisy_val_loader)}/{len(self.noisy_test_loader)} batches")
        if self.TRACK_BIAS:
            n = len(self.noisy_student_loader.dataset)
            if n > 50:
                self.tracked_sample_indices = np.random.choice(n, 50, replace=False).tolist()  # I/train.py:278-285

    def _setup_model(self):
        self.model = SSRLModel(cfg=self.cfg).to(self.device)
        if getattr(self.cfg, "PRINT_MODEL_INFO", False):
            total = sum(p.numel() for p in self.model.parameters())
            logger.info(f"model parameters: {total:,} total, {sum(p.numel() for p in self.model.parameters() if p.requires_grad):,} trainable")

    def _load_pretrained_weights(self):
        path = getattr(self.cfg, self.WEIGHT_FIELD, None)
        if not path or not os.path.exists(path):
            if self._require_pretrained:
                raise FileNotFoundError(f"pretrained weight file missing: {path}")  # I/train.py:304-306
            logger.warning(f"pretrained weight file missing ({path}): continuing from the current initialisation")
        else:
            self.model.load_complete_pretrained_weights(path)
        self.model._init_teacher_network()

    def _batch_to_device(self, batch):
        ni = batch["net_input"]
        return ni["feats"].to(self.device, non_blocking=True), ni["padding_mask"].to(self.device, non_blocking=True)

    def _run_anchor_calibration(self):
        """One-off per-class anchors from clean vs noisy validation certainty (I/train.py:317-357):
        anchor_c = clamp(mu_clean_c - K*sigma_clean_c, 0) * (mean noisy score_c / mean clean score_c)."""
        cfg = self.cfg
        self.model.eval()
        per = {"clean": [[] for _ in range(self.num_classes)], "noisy": [[] for _ in range(self.num_classes)]}
        with torch.no_grad():
            for name, loader in (("clean", self.clean_train_loader), ("noisy", self.noisy_val_loader)):
                for batch in loader:
                    feats, mask = self._batch_to_device(batch)
                    probs = F.softmax(self.model.predict(feats, mask), dim=1)
                    scores, _ = DACPManager.calculate_certainty_scores(probs)
                    for s, y in zip(scores.tolist(), batch["labels"].tolist()):
                        per[name][y].append(s)
        dev = self.device
        mean = lambda ls: torch.tensor([np.mean(s) if s else 0 for s in ls], device=dev)
        mu_c, mu_n = mean(per["clean"]), mean(per["noisy"])
        sd_c = torch.tensor([np.std(s) if s else 0 for s in per["clean"]], device=dev)
        self.calibrated_anchors = (torch.clamp(mu_c - cfg.ANCHOR_STD_K * sd_c, min=0) * (mu_n / (mu_c + 1e-8))).float()
        logger.info("calibrated anchors: " + ", ".join(f"{n}={a:.4f}" for n, a in zip(self.class_names, self.calibrated_anchors.tolist())))
        self.model.train()

    def _capacity(self):
        """(B, T) capacity of the engine: the batch size and the longest utterance of any training split."""
        T = 32
        for ld in (self.clean_train_loader, self.noisy_student_loader):
            if hasattr(ld, "sizes") and len(ld.sizes):
                T = max(T, int(ld.sizes.max()))
            else:
                for b in ld:
                    T = max(T, int(b["net_input"]["feats"].shape[1]))
        bk = int(getattr(self.cfg, "DAD_T_BUCKET", 0))
        if bk > 0:
            T = (T + bk - 1) // bk * bk  # bucketed padding (dataload.UtteranceLoader.t_bucket) rounds the longest batch up
        return int(self.cfg.BATCH_SIZE), T

    def _setup_training_components(self):
        cfg = self.cfg
        B, T = self._capacity()
        mode = _lib.GEMM_TC if str(getattr(cfg, "DAD_GEMM_MODE", "tc")).lower() in ("tc", "f16x3") else _lib.GEMM_FP32
        self.engine = DADStepEngine(cfg, B, T, device=self.device, dataset=self.DATASET, gemm_mode=mode,
                                    use_graph=bool(getattr(cfg, "DAD_USE_CUDA_GRAPH", True)), noise_mode="philox",
                                    seed=int(getattr(cfg, "RANDOM_SEED", 42)))
        self.engine.load_state_dict(self.model.state_dict())
        # HBM-resident training splits assemble every batch straight into the engine's input buffers (no copy in set_inputs)
        for ld, dom in ((self.clean_train_loader, "clean"), (self.noisy_student_loader, "noisy")):
            if hasattr(ld, "sink"):
                ld.sink = (lambda b_, t_, d_=dom: self.engine.input_views(d_, b_, t_))
        if self.calibrated_anchors is not None:
            self.engine.anchors.copy_(self.calibrated_anchors)
        else:
            # upstream passes calibrated_anchors=None into calculate_mask in this configuration and crashes on `None.to(...)`
            # (I/utils.py:492); the intended meaning -- no floor under the dynamic thresholds -- is what runs here
            self.engine.anchors.fill_(-1e30)
        self.optimizer = EngineOptimizer(self.engine, cfg)
        self.scheduler = CosineScheduler(self.optimizer, cfg.LEARNING_RATE, cfg.EPOCHS) if cfg.LEARNING_RATE_SCHEDULER == "cosine" else None
        self.ce_criterion = torch.nn.CrossEntropyLoss(label_smoothing=cfg.LABEL_SMOOTHING_FACTOR if cfg.USE_LABEL_SMOOTHING else 0.0)
        self.kl_criterion = torch.nn.KLDivLoss(reduction="none")
        # DACP state is the engine's device state: the manager object the reference exposes shares those tensors
        self.dacp_manager = DACPManager(self.num_classes, cfg.EPOCHS, self.device)
        self.dacp_manager.class_quality_scores = self.engine.quality
        self.dacp_manager.ema_thresholds = self.engine.ema_thr
        self.ecda_criterion = ECDALoss()

    def _setup_augmentation(self):
        self.augmenter = DataAugmentation()

    # ------------------------------------------------------------------ schedule
    def is_warmup_phase(self, epoch):
        return epoch < self.WARMUP_EPOCHS

    def update_loss_weights(self, epoch):
        """I/train.py:380-395 (the engine derives the same weights from cfg for the kernels; these attributes mirror them)."""
        sc = self.engine.scalars(epoch)
        self.current_consistency_weight, self.weight_ecda = sc.w_kd, sc.w_ecda

    # ------------------------------------------------------------------ the hot loop
    def train_step(self, clean_batch, noisy_batch, epoch):
        """One FULL optimisation step (see the module docstring).  Returns detached device scalars."""
        eng = self.engine
        eng.set_inputs(clean_batch, None if self.is_warmup_phase(epoch) else noisy_batch)
        eng.step(epoch)
        if self.tracked_sample_indices and not self.is_warmup_phase(epoch) and isinstance(noisy_batch.get("id"), torch.Tensor):
            gv = _lib.pack_global_views(eng.pack, eng.B, 1)
            n = eng.n_noisy
            self._bias_pending.append((epoch, noisy_batch["id"].clone(), gv["pred"][0][:n].clone(), gv["score"][0][:n].clone(),
                                       eng.mask[:n].clone()))  # device clones, resolved at epoch end: no sync here
        ll = eng.loss_last.clone()
        out = {"total_loss": ll[0], "supervised_ce_loss": ll[1], "consistency_loss": ll[2], "ecda_loss": ll[3]}
        if "scl_loss" in self.EXTRA_LOSS_KEYS or eng.scalars(epoch).w_scl > 0:
            out["scl_loss"] = ll[4]
        return out

    def train_epoch(self, epoch):
        self.model.train()
        self.update_loss_weights(epoch)
        eng = self.engine
        eng.set_lr(self.optimizer.param_groups[0]["lr"])
        eng.loss_acc.zero_()
        clean_iter, noisy_iter = iter(self.clean_train_loader), iter(self.noisy_student_loader)
        n_batches = min(len(self.clean_train_loader), len(self.noisy_student_loader))  # I/train.py:479-483
        for _ in range(n_batches):
            self.train_step(next(clean_iter), next(noisy_iter), epoch)
        avg = eng.pop_epoch_losses()  # the epoch's single loss read-back
        if "scl_loss" not in self.EXTRA_LOSS_KEYS and avg["scl_loss"] == 0.0:
            del avg["scl_loss"]  # I/train.py returns no 'scl_loss' key; C/ and E/ return an always-zero one
        if not self.is_warmup_phase(epoch):
            eng.end_epoch(epoch)  # DACPManager.update_class_quality_scores_epoch from device accumulators (I/train.py:499)
            self._flush_bias_log()
            if self.TRACK_BIAS:
                q = eng.quality
                self.training_history["dacp_ema_thresholds"].append(eng.ema_thr.tolist())
                self.training_history["dacp_class_quality"].append(q.tolist())
                self.training_history["ecda_class_attention"].append(
                    torch.exp(self.cfg.ECDA_CLASS_ATTENTION_LAMBDA * (q.mean() - q)).tolist())
        if self.scheduler:
            self.scheduler.step()
        return avg

    def _flush_bias_log(self):
        tracked = set(self.tracked_sample_indices or ())
        for epoch, ids, pred, score, mask in self._bias_pending:
            for i, p, s, m in zip(ids.tolist(), pred.tolist(), score.tolist(), mask.tolist()):
                if int(i) in tracked:
                    self.bias_analysis_log.append({"epoch": epoch, "sample_id": int(i), "pseudo_label": int(p),
                                                   "certainty_score": float(s), "is_masked_in": bool(m)})
        self._bias_pending = []

    # ------------------------------------------------------------------ evaluation
    def _sync_model_from_engine(self):
        self.model.load_state_dict(self.engine.state_dict())

    def _predict_all(self, data_loader, use_teacher):
        preds, labels = [], []
        with torch.no_grad():
            for batch in data_loader:
                feats, mask = self._batch_to_device(batch)
                preds.append(self.model.predict(feats, mask, use_teacher=use_teacher).argmax(dim=1))
                labels.append(batch["labels"])
        return torch.cat(preds).cpu().numpy(), torch.cat(labels).cpu().numpy()

    def _predict_all_both(self, data_loader):
        """student AND teacher predictions of a whole split from one pass over its batches (SSRLModel.predict_both)."""
        ps, pt, labels = [], [], []
        with torch.no_grad():
            for batch in data_loader:
                feats, mask = self._batch_to_device(batch)
                ls, lt = self.model.predict_both(feats, mask)
                ps.append(ls.argmax(dim=1)); pt.append(lt.argmax(dim=1))
                labels.append(batch["labels"])
        return torch.cat(ps).cpu().numpy(), torch.cat(pt).cpu().numpy(), torch.cat(labels).cpu().numpy()

    def validate(self, data_loader, domain_name="unknown"):
        """Accuracy, balanced ('weighted') accuracy, F1s, per-class stats, confusion matrix (I/train.py:522-564)."""
        from sklearn.metrics import accuracy_score, balanced_accuracy_score, confusion_matrix, f1_score, precision_recall_fscore_support
        if getattr(self, "_eval_from_engine", True):
            self._sync_model_from_engine()
        self.model.eval()
        if "noisy" in domain_name.lower() and not self.is_warmup_phase(getattr(self, "current_epoch", 0)):
            # I/train.py:530-540 runs the loader twice (student, then teacher); here both networks see each batch in one pass
            p_student, p_teacher, y = self._predict_all_both(data_loader)
            self.training_history[f"disagreement_rate_{domain_name.lower()}"].append(float(np.mean(p_student != p_teacher)))
        else:
            p_student, y = self._predict_all(data_loader, use_teacher=False)
        lab = list(range(self.num_classes))
        cm = confusion_matrix(y, p_student, labels=lab)
        prec, rec, f1, sup = precision_recall_fscore_support(y, p_student, average=None, zero_division=0, labels=lab)
        return {"accuracy": accuracy_score(y, p_student) * 100, "weighted_accuracy": balanced_accuracy_score(y, p_student) * 100,
                "f1_weighted": f1_score(y, p_student, average="weighted", zero_division=0) * 100,
                "f1_macro": f1_score(y, p_student, average="macro", zero_division=0) * 100,
                "precision_per_class": prec.tolist(), "recall_per_class": rec.tolist(), "f1_per_class": f1.tolist(),
                "support_per_class": sup.tolist(), "confusion_matrix": cm}

    def check_early_stopping(self, noisy_results, is_best):
        if not self.cfg.EARLY_STOPPING:
            return False
        if is_best:
            self.patience_counter = 0
            return False
        self.patience_counter += 1
        return self.patience_counter >= self.cfg.PATIENCE

    # ------------------------------------------------------------------ persistence
    def save_checkpoint(self, epoch, clean_results, noisy_results, is_best=False):
        """{'epoch','model_state_dict','optimizer_state_dict','clean_results','noisy_results'} (I/train.py:581-592) + the DACP
        state the reference forgets ('dacp_state'), so a run can be resumed."""
        if not is_best:
            return
        self._sync_model_from_engine()
        ckpt = {"epoch": epoch, "model_state_dict": self.model.state_dict(), "optimizer_state_dict": self.optimizer.state_dict(),
                "clean_results": clean_results, "noisy_results": noisy_results,
                "dacp_state": {"ema_thresholds": self.engine.ema_thr.cpu(), "class_quality_scores": self.engine.quality.cpu(),
                               "calibrated_anchors": self.engine.anchors.cpu()}}
        self.best_results.update({"epoch": epoch, "clean_results": clean_results, "noisy_results": noisy_results})
        path = os.path.join(self.results_dir, "models", self.BEST_NAME)
        torch.save(ckpt, path)
        self._save_confusion_matrices(clean_results, noisy_results, epoch, is_best_result=True)
        self._save_detailed_results(clean_results, noisy_results, epoch, is_best_result=True)
        logger.info(f"best model saved: {path}")

    def load_checkpoint(self, path):
        """Resume from a best-model file written by save_checkpoint (or by the reference: DACP state then starts fresh)."""
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ckpt["model_state_dict"])
        self.engine.load_state_dict(ckpt["model_state_dict"], reset_optimizer=True)
        if "optimizer_state_dict" in ckpt and ckpt["optimizer_state_dict"].get("state"):
            self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        ds = ckpt.get("dacp_state")
        if ds:
            self.engine.ema_thr.copy_(ds["ema_thresholds"]); self.engine.quality.copy_(ds["class_quality_scores"])
            self.engine.anchors.copy_(ds["calibrated_anchors"])
        return ckpt["epoch"]

    def _save_confusion_matrices(self, clean_results, noisy_results, epoch, is_best_result=False):
        """PNG heat-maps as upstream (I/train.py:594-606) when matplotlib + seaborn are importable; skipped otherwise."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            import seaborn as sns
        except Exception:  # noqa: BLE001 -- plotting is optional tooling, not part of the training path
            return
        fig, (ax1, ax2) = plt.subplots(1, 2, figsize=(16, 7))
        for ax, res, cmap, title in ((ax1, clean_results, "Blues", "Clean Domain"), (ax2, noisy_results, "Oranges", f"Noisy Domain ({self.noise_info['display_name']})")):
            sns.heatmap(res["confusion_matrix"], annot=True, fmt="d", cmap=cmap, xticklabels=self.class_names, yticklabels=self.class_names, ax=ax)
            ax.set_title(f"{title} (Epoch {epoch + 1})\nAcc: {res['accuracy']:.2f}%, W-Acc: {res['weighted_accuracy']:.2f}%")
        if is_best_result:
            fig.suptitle("BEST RESULTS", fontsize=16, weight="bold")
        name = f"BEST_confusion_matrices_epoch_{epoch + 1}.png" if is_best_result else f"confusion_matrices_epoch_{epoch + 1}.png"
        plt.savefig(os.path.join(self.results_dir, "plots", name), dpi=300)
        plt.close()

    def _save_detailed_results(self, clean_results, noisy_results, epoch, is_best_result=False):
        """JSON report in the reference's layout (I/train.py:608-633)."""
        pca = lambda cm: [float(cm[i, i] / cm[i, :].sum()) if cm[i, :].sum() > 0 else 0.0 for i in range(len(cm))]
        summ = lambda r: {"acc": f"{r['accuracy']:.2f}%", "w_acc": f"{r['weighted_accuracy']:.2f}%", "w_f1": f"{r['f1_weighted']:.2f}%"}
        det = lambda r: {"precision": r["precision_per_class"], "recall": r["recall_per_class"], "f1": r["f1_per_class"],
                         "support": r["support_per_class"], "accuracy": pca(r["confusion_matrix"])}
        out = {"info": {"noise_config": self.noise_info, "fold": self.fold + 1, "epoch": epoch + 1, "is_best": is_best_result},
               "summary": {"clean": summ(clean_results), "noisy": summ(noisy_results)},
               "details": {"class_names": self.class_names, "clean": det(clean_results), "noisy": det(noisy_results)}}
        name = f"BEST_detailed_results_epoch_{epoch + 1}.json" if is_best_result else f"detailed_results_epoch_{epoch + 1}.json"
        with open(os.path.join(self.results_dir, "reports", name), "w", encoding="utf-8") as f:
            json.dump(out, f, indent=4, ensure_ascii=False)

    # ------------------------------------------------------------------ main loop
    def train(self):
        cfg = self.cfg
        for epoch in range(cfg.EPOCHS):
            avg = self.train_epoch(epoch)
            for k, v in avg.items():
                self.training_history[k].append(v)
            if (epoch + 1) % cfg.VALIDATION_INTERVAL == 0 or not self.is_warmup_phase(epoch):
                self.current_epoch = epoch
                clean_results = self.validate(self.clean_val_loader, "Clean")
                noisy_results = self.validate(self.noisy_val_loader, "Noisy")
                is_best = noisy_results["weighted_accuracy"] > self.best_noisy_weighted_acc + cfg.MIN_DELTA
                if is_best:
                    self.best_noisy_acc, self.best_clean_acc = noisy_results["accuracy"], clean_results["accuracy"]
                    self.best_noisy_weighted_f1, self.best_clean_weighted_f1 = noisy_results["f1_weighted"], clean_results["f1_weighted"]
                    self.best_noisy_weighted_acc, self.best_clean_weighted_acc = noisy_results["weighted_accuracy"], clean_results["weighted_accuracy"]
                self.save_checkpoint(epoch, clean_results, noisy_results, is_best)
                logger.info(f"Epoch {epoch + 1}/{cfg.EPOCHS} | Losses: Total={avg.get('total_loss', 0):.4f} CE={avg.get('supervised_ce_loss', 0):.4f} "
                            f"KD={avg.get('consistency_loss', 0):.4f} ECDA={avg.get('ecda_loss', 0):.4f} | Noisy WA: {noisy_results['weighted_accuracy']:.2f}%"
                            f"{' *' if is_best else ''}")
                if self.check_early_stopping(noisy_results, is_best):
                    logger.info("early stopping")
                    break
        self._save_analysis_data()
        self._evaluate_on_test_set()
        return {"best_noisy_weighted_acc": self.best_noisy_weighted_acc, "results_dir": self.results_dir}

    def _save_analysis_data(self):
        """training_history.json (+ confirmation_bias_log.json), I/train.py:674-700."""
        hist = {k: [i.tolist() if isinstance(i, np.ndarray) else i for i in v] for k, v in self.training_history.items()}
        with open(os.path.join(self.results_dir, "reports", "training_history.json"), "w") as f:
            json.dump(hist, f, indent=4)
        if self.bias_analysis_log:
            with open(os.path.join(self.results_dir, "reports", "confirmation_bias_log.json"), "w") as f:
                json.dump(self.bias_analysis_log, f, indent=4)

    def _evaluate_on_test_set(self):
        """Reload the best checkpoint and evaluate both test splits (I/train.py:702-762)."""
        path = os.path.join(self.results_dir, "models", self.BEST_NAME)
        if not os.path.exists(path):
            logger.warning(f"best model not found ({path}); skipping the test-set evaluation")
            return None
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ckpt["model_state_dict"])
        self._eval_from_engine = False  # evaluate the checkpoint, not the engine's current weights
        try:
            clean_t = self.validate(self.clean_test_loader, "Clean_Test")
            noisy_t = self.validate(self.noisy_test_loader, "Noisy_Test")
        finally:
            self._eval_from_engine = True
        self._save_confusion_matrices(clean_t, noisy_t, 999, is_best_result=False)
        self._save_detailed_results(clean_t, noisy_t, 999, is_best_result=False)
        fmt = lambda r: {"accuracy": f"{r['accuracy']:.2f}%", "weighted_accuracy": f"{r['weighted_accuracy']:.2f}%", "weighted_f1": f"{r['f1_weighted']:.2f}%"}
        rep = {"info": {"noise_config": self.noise_info, "fold": self.fold + 1, "evaluation_type": "Final Test Set Evaluation",
                        "timestamp": datetime.now().isoformat()},
               "final_test_results": {"clean_domain": fmt(clean_t), "noisy_domain": fmt(noisy_t)},
               "comparison_with_validation": {"validation_best_noisy_weighted_acc": f"{self.best_noisy_weighted_acc:.2f}%",
                                              "test_noisy_weighted_acc": f"{noisy_t['weighted_accuracy']:.2f}%",
                                              "performance_gap": f"{(noisy_t['weighted_accuracy'] - self.best_noisy_weighted_acc):.2f}%"}}
        with open(os.path.join(self.results_dir, "reports", "FINAL_test_set_results.json"), "w", encoding="utf-8") as f:
            json.dump(rep, f, indent=4, ensure_ascii=False)
        return clean_t, noisy_t


def run_main(trainer_cls, default_fold, argv=None):
    """`python train*.py [--fold k]` -- the CLI the READMEs promise (README.md:140,155,170) and upstream never parses;
    the default reproduces the hard-coded behaviour (I/train.py:778, C/train_CASIA.py:721, E/train_emodb.py:672-675)."""
    import argparse
    ap = argparse.ArgumentParser(description=f"{trainer_cls.DISPLAY} cross-domain DAD training (DACP + ECDA) on B200")
    ap.add_argument("--fold", type=int, default=None, help=f"1-based fold (1..{trainer_cls.N_FOLDS_TOTAL}); default = the reference's hard-coded fold")
    ap.add_argument("--experiment-name", default=None)
    ap.add_argument("--epochs", type=int, default=None, help="override cfg.EPOCHS")
    ap.add_argument("--synthetic", action="store_true", help="train on synthetic stand-in features (no dataset files needed)")
    args = ap.parse_args(argv)
    cfg = trainer_cls.cfg
    if args.epochs is not None:
        cfg.EPOCHS = args.epochs
    fold = default_fold if args.fold is None else args.fold - 1
    dev = "cuda" if (getattr(cfg, "DAD_DEVICE_DATA", True) and torch.cuda.is_available()) else None
    loaders = trainer_cls.synthetic_loaders(fold, device=dev) if args.synthetic else None
    try:
        trainer = trainer_cls(fold=fold, experiment_name=args.experiment_name, loaders=loaders, require_pretrained=not args.synthetic)
        res = trainer.train()
        logger.info(f"done: best noisy weighted accuracy {res['best_noisy_weighted_acc']:.2f}% -> {res['results_dir']}")
        return res
    except Exception as e:  # noqa: BLE001 -- reference behaviour: log and continue (I/train.py:775-789)
        logger.error(f"fold {fold + 1} failed: {e}", exc_info=True)
        return None
