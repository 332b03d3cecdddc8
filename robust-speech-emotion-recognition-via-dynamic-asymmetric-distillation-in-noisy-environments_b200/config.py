"""IEMOCAP cross-domain DAD configuration -- field names and values of I/config.py (drop-in)."""
import os

from ._config_base import data_root, populate

populate(globals(), "iemocap", dict(
    _WEIGHT_KEY="PRETRAINED_IEMOCAP_WEIGHT",
    CLEAN_DATA_DIR=os.path.join(data_root(), "processed_features_IEMOCAP"),
    NOISY_DATA_DIR=os.path.join(data_root(), "processed_features_IEMOCAP_noisy", "root1-babble-0db"),
    SESSION_SAMPLES=[1085, 1023, 1151, 1031, 1241],
    LABEL_DICT={"ang": 0, "hap": 1, "neu": 2, "sad": 3}, CLASS_NAMES=["ang", "hap", "neu", "sad"],
    PRETRAINED_IEMOCAP_WEIGHT=os.path.join(data_root(), "pretrained", "iemocap_best_model_fold_2.ckpt"),
    N_FOLDS=2,
))
