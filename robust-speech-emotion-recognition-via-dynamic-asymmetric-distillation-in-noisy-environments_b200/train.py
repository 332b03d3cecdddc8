"""IEMOCAP cross-domain DAD training -- drop-in for I/train.py (IEMOCAPCrossDomainTrainer, main); `python train.py [--fold k]`."""
from . import config as cfg
from . import dataload
from .trainer import CrossDomainTrainer, run_main


class IEMOCAPCrossDomainTrainer(CrossDomainTrainer):
    """I/train.py:44-803.  5 session folds (train 3 / val 1 / test 1), confirmation-bias + DACP/ECDA history taps."""
    DATASET, DISPLAY = "iemocap", "IEMOCAP"
    RESULTS_BASE, BEST_NAME = "iemocap_mutil-noisy_cross_domain_results", "iemocap_cross_domain_best.pth"
    WEIGHT_FIELD, N_FOLDS_TOTAL, TRACK_BIAS = "PRETRAINED_IEMOCAP_WEIGHT", 5, True
    cfg = cfg

    def _make_loaders(self, batch_size):
        fold_id = self.fold + 1
        dev = self._data_device()
        clean = dataload.get_cv_dataloaders(cfg.CLEAN_DATA_DIR, batch_size, fold_id=fold_id, label_dict=cfg.LABEL_DICT, device=dev)
        noisy = dataload.get_cv_dataloaders_noisy(cfg.NOISY_DATA_DIR, batch_size, fold_id=fold_id, label_dict=cfg.LABEL_DICT, device=dev)
        return clean, noisy

    @classmethod
    def synthetic_loaders(cls, fold, utt_per_session=96, seed=0, device=None):
        """Loader factory over synthetic stand-in features with IEMOCAP's 5-session structure (offline runs / tests)."""
        fs_c, fs_n = dataload.synthetic_feature_sets(5, utt_per_session, seed=seed)

        def make(batch_size):
            return (dataload.get_cv_dataloaders(None, batch_size, fold + 1, cfg.LABEL_DICT, feature_set=fs_c, device=device),
                    dataload.get_cv_dataloaders_noisy(None, batch_size, fold + 1, cfg.LABEL_DICT, feature_set=fs_n, device=device))
        return make


def main(argv=None):
    """Upstream main() trains the single fold cfg.N_FOLDS-1 (I/train.py:765-803); --fold k (1-based) selects another."""
    return run_main(IEMOCAPCrossDomainTrainer, default_fold=cfg.N_FOLDS - 1, argv=argv)


if __name__ == "__main__":
    main()
