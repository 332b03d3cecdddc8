"""EMODB cross-domain DAD training -- drop-in for E/train_emodb.py (FixedEMODBCrossDomainTrainer); `python train_emodb.py [--fold k]`."""
from . import config_emodb as cfg
from .train_CASIA import FixedCASIACrossDomainTrainer
from .trainer import run_main


class FixedEMODBCrossDomainTrainer(FixedCASIACrossDomainTrainer):
    """E/train_emodb.py:44-700.  10 speaker folds; no USE_DACP / USE_ECDA switches -- always DACP, ECDA whenever its weight is
    positive (E/train_emodb.py:418-437); 'scl_loss' key as CASIA."""
    DATASET, DISPLAY = "emodb", "EMODB"
    RESULTS_BASE, BEST_NAME = "emodb_mutil-noisy_cross_domain_results", "emodb_cross_domain_best.pth"
    WEIGHT_FIELD, N_FOLDS_TOTAL = "PRETRAINED_EMODB_WEIGHT", 10
    cfg = cfg
    EMODB = True


def main(argv=None):
    """Upstream hard-codes fold = 3 (E/train_emodb.py:672-675); --fold k (1-based) selects another."""
    return run_main(FixedEMODBCrossDomainTrainer, default_fold=3, argv=argv)


if __name__ == "__main__":
    main()
