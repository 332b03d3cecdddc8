"""EMODB cross-domain DAD configuration -- field names and values of E/config_emodb.py (drop-in)."""
import os

from ._config_base import data_root, populate

_clean = os.path.join(data_root(), "processed_features_EMODB", "processed_features_clean")
_noisy = os.path.join(data_root(), "processed_features_EMODB_noisy", "root1-babble-10db")
populate(globals(), "emodb", dict(
    _WEIGHT_KEY="PRETRAINED_EMODB_WEIGHT",
    CLEAN_DATA_DIR=_clean, NOISY_DATA_DIR=_noisy, CLEAN_FEAT_PATH=_clean + "/train", NOISY_FEAT_PATH=_noisy + "/train",
    LABEL_DICT={"angry": 0, "happy": 1, "neutral": 2, "sad": 3}, CLASS_NAMES=["angry", "happy", "neutral", "sad"],
    PRETRAINED_EMODB_WEIGHT=os.path.join(data_root(), "pretrained", "emodb_best_model_fold_4.ckpt"),
    DACP_CALIBRATION_STRENGTH_LAMBDA=0.3, DACP_QUALITY_SMOOTHING_BETA=0.8, FIXED_CONFIDENCE_THRESHOLD=0.75,
    LEARNING_RATE=5e-3, WEIGHT_ECDA=0.1,
))
