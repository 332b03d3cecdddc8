"""Shared defaults for the three dataset config modules (config.py / config_casia.py / config_emodb.py).

The reference keeps its hyper-parameters as module-level constants that runners override with setattr after an
importlib.reload (I/run_granular_ablations.py:25-44); every field name below is part of that drop-in surface
(values from I/config.py:24-151).  `populate(globals(), ...)` writes them into a config module."""
import os

import torch

_COMMON = dict(
    # basics (I/config.py:24-29)
    RANDOM_SEED=42, USE_CUDA=True, GRADIENT_CLIPPING=True, MAX_GRAD_NORM=1.0,
    BATCH_SIZE=64, NUM_WORKERS=0,
    # model (I/config.py:54-60)
    INPUT_DIM=768, HIDDEN_DIM=256, DROPOUT_RATE=0.1, EMA_MOMENTUM=0.995,
    # phases (I/config.py:63-66)
    WARMUP_EPOCHS=30, SCL_START_EPOCH=5001, ECDA_START_EPOCH=30, DISABLE_MMD=True,
    # DACP (I/config.py:70-79)
    DACP_QUALITY_SMOOTHING_BETA=0.9, DACP_SENSITIVITY_K=10.0, DACP_QUANTILE_START=0.4, DACP_QUANTILE_END=0.80,
    DACP_CALIBRATION_STRENGTH_LAMBDA=0.9, DACP_THRESHOLD_SMOOTHING_ALPHA=0.9,
    # ablation switches (I/config.py:82-92)
    USE_DACP=True, USE_ECDA=True, FIXED_CONFIDENCE_THRESHOLD=0.9, USE_ENTROPY_IN_SCORE=True, USE_CLASS_AWARE_MMD=True,
    ANCHOR_CALIBRATION_ENABLED=True, ANCHOR_STD_K=1.5,
    # ECDA (I/config.py:96-105)
    ECDA_CLASS_ATTENTION_LAMBDA=1.0, ECDA_COMPACTNESS_WEIGHT_GAMMA=0.1, ECDA_REPULSION_WEIGHT_DELTA=0.1,
    WEIGHT_CONSISTENCY=1.0, TARGET_SCL_WEIGHT=0.0, WEIGHT_ECDA=0.3,
    # optimisation (I/config.py:108-116)
    EPOCHS=500, LEARNING_RATE=5e-4, WEIGHT_DECAY=1e-5, LEARNING_RATE_SCHEDULER="cosine",
    USE_LABEL_SMOOTHING=True, LABEL_SMOOTHING_FACTOR=0.05,
    # augmentation (I/config.py:120-128)
    WEAK_NOISE_STD=0.01, STRONG_NOISE_STD=0.05, TEMPORAL_MASK_RATIO=0.1,
    # validation / saving (I/config.py:131-137)
    VALIDATION_INTERVAL=5, PLOT_INTERVAL=10, SAVE_BEST_MODEL=True, PLOT_CONFUSION_MATRIX=True,
    # progressive weights + early stopping (I/config.py:140-148)
    PROGRESSIVE_TRAINING=True, INITIAL_CONSISTENCY_WEIGHT=0.1, FINAL_CONSISTENCY_WEIGHT=0.3, WEIGHT_RAMP_EPOCHS=30,
    EARLY_STOPPING=True, PATIENCE=50, MIN_DELTA=0.001,
    PRINT_MODEL_INFO=True, DEBUG_MODE=False, VERBOSE_LOGGING=True,
    # B200 additions (not in the reference): kernel arithmetic + noise source of the fused step
    DAD_GEMM_MODE=os.environ.get("DAD_GEMM", "tc"), DAD_USE_CUDA_GRAPH=True,
    DAD_T_BUCKET=0,  # > 0: training batches are padded to a multiple of this many frames so that step shapes repeat (CUDA-graph replay); 0 = reference padding
    DAD_DEVICE_DATA=True,  # keep each split's features resident in HBM and assemble batches on the device (dataload.DeviceUtteranceLoader)
    DAD_SYNTHETIC_DATA=False,  # trainers built without explicit loaders use synthetic stand-in features (offline runs, runner tests)
)


def data_root():
    return os.environ.get("DAD_DATA_ROOT", os.path.join(os.path.expanduser("~"), "DATA"))


def populate(ns, tag, overrides):
    ns.update(_COMMON)
    ns.update(overrides)
    ns["DEVICE"] = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    ns["NUM_CLASSES"] = len(ns["LABEL_DICT"])
    ns["AUGMENT_DROPOUT_RATE"] = ns["FEATURE_DROPOUT_RATE"] = ns["DROPOUT_RATE"]
    ns["NOISE_INJECTION_STD"] = ns["STRONG_NOISE_STD"]
    ns["MODEL_SAVE_DIR"], ns["LOG_DIR"], ns["RESULTS_DIR"] = (f"{tag}_cross_domain_models", f"{tag}_cross_domain_logs",
                                                             f"{tag}_cross_domain_results")

    def validate_config():
        """Same checks as I/config.py:156-178 (asserts on sizes, warnings on missing paths)."""
        for key in ("CLEAN_DATA_DIR", "NOISY_DATA_DIR"):
            if not os.path.exists(os.path.dirname(ns[key])):
                print(f"warning: {key} parent directory missing: {ns[key]}")
        w = ns.get(ns["_WEIGHT_KEY"])
        if w and not os.path.exists(w):
            print(f"warning: pretrained weight missing: {w}")
        assert ns["BATCH_SIZE"] > 0 and ns["EPOCHS"] > 0 and ns["LEARNING_RATE"] > 0
        assert ns["NUM_CLASSES"] == 4 and len(ns["LABEL_DICT"]) == ns["NUM_CLASSES"]

    def print_config():
        for k in ("DEVICE", "BATCH_SIZE", "EPOCHS", "WARMUP_EPOCHS", "LEARNING_RATE", "EMA_MOMENTUM", "WEIGHT_ECDA",
                  "CLEAN_DATA_DIR", "NOISY_DATA_DIR", ns["_WEIGHT_KEY"], "MODEL_SAVE_DIR"):
            print(f"   {k}: {ns[k]}")

    def setup_environment():
        """Create output dirs and seed python / numpy / torch (I/config.py:197-218)."""
        import random

        import numpy as np
        for d in (ns["MODEL_SAVE_DIR"], ns["LOG_DIR"], ns["RESULTS_DIR"]):
            os.makedirs(d, exist_ok=True)
        random.seed(ns["RANDOM_SEED"])
        np.random.seed(ns["RANDOM_SEED"])
        torch.manual_seed(ns["RANDOM_SEED"])
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(ns["RANDOM_SEED"])

    ns.update(validate_config=validate_config, print_config=print_config, setup_environment=setup_environment)
