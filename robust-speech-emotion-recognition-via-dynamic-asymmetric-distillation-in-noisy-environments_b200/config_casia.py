"""CASIA cross-domain DAD configuration -- field names and values of C/config_casia.py (drop-in)."""
import os

from ._config_base import data_root, populate

_clean = os.path.join(data_root(), "fix_CASIA", "processed_features_clean")
_noisy = os.path.join(data_root(), "processed_features_CASIA_real_noise", "processed_features_babble_10db")
populate(globals(), "casia", dict(
    _WEIGHT_KEY="PRETRAINED_CASIA_WEIGHT",
    CLEAN_DATA_DIR=_clean, NOISY_DATA_DIR=_noisy, CLEAN_FEAT_PATH=_clean + "/train", NOISY_FEAT_PATH=_noisy + "/train",
    LABEL_DICT={"angry": 0, "happy": 1, "neutral": 2, "sad": 3}, CLASS_NAMES=["angry", "happy", "neutral", "sad"],
    PRETRAINED_CASIA_WEIGHT=os.path.join(data_root(), "pretrained", "casia_best_model_fold_4.ckpt"),
    DACP_CALIBRATION_STRENGTH_LAMBDA=0.1, ECDA_COMPACTNESS_WEIGHT_GAMMA=0.05, ECDA_REPULSION_WEIGHT_DELTA=0.05,
    FIXED_CONFIDENCE_THRESHOLD=0.75, USE_DACP=False, USE_ECDA=False, WEIGHT_ECDA=0.35,
))
