"""CASIA cross-domain DAD training -- drop-in for C/train_CASIA.py (FixedCASIACrossDomainTrainer); `python train_CASIA.py [--fold k]`."""
from . import config_casia as cfg
from . import dataload
from .trainer import CrossDomainTrainer, run_main


class FixedCASIACrossDomainTrainer(CrossDomainTrainer):
    """C/train_CASIA.py:44-760.  4 speaker folds; cfg.USE_DACP / USE_ECDA default False (fixed 0.75 threshold, no ECDA);
    train_step returns the extra always-zero 'scl_loss' (C/train_CASIA.py:442,503-504)."""
    DATASET, DISPLAY = "casia", "CASIA"
    RESULTS_BASE, BEST_NAME = "casia_mutil-noise_cross_domain_results", "casia_cross_domain_best.pth"
    WEIGHT_FIELD, N_FOLDS_TOTAL, EXTRA_LOSS_KEYS, REAL_NOISE_DIRS = "PRETRAINED_CASIA_WEIGHT", 4, ("scl_loss",), True
    cfg = cfg
    EMODB = False

    def __init__(self, fold=3, experiment_name=None, **kw):
        super().__init__(fold=fold, experiment_name=experiment_name, **kw)

    def _make_loaders(self, batch_size):
        dev = self._data_device()
        tr, va, te, ncls, names = dataload.create_speaker_isolated_loaders(self.cfg.CLEAN_FEAT_PATH, batch_size, self.fold,
                                                                          self.cfg.LABEL_DICT, noisy=False, emodb=self.EMODB, device=dev)
        noisy = dataload.create_speaker_isolated_loaders(self.cfg.NOISY_FEAT_PATH, batch_size, self.fold, self.cfg.LABEL_DICT,
                                                         noisy=True, emodb=self.EMODB, device=dev)
        return (tr, va, te, names, ncls), noisy

    @classmethod
    def synthetic_loaders(cls, fold, utt_per_speaker=96, seed=0, device=None):
        names = [f"spk_{i:02d}" for i in range(cls.N_FOLDS_TOTAL)]
        fs_c, fs_n = dataload.synthetic_feature_sets(cls.N_FOLDS_TOTAL, utt_per_speaker, seed=seed, group_names=names)
        ld = cls.cfg.LABEL_DICT

        def make(batch_size):
            tr, va, te, ncls, cn = dataload.create_speaker_isolated_loaders(None, batch_size, fold, ld, False, cls.EMODB, feature_set=fs_c, device=device)
            return (tr, va, te, cn, ncls), dataload.create_speaker_isolated_loaders(None, batch_size, fold, ld, True, cls.EMODB, feature_set=fs_n,
                                                                                   device=device)
        return make


def main(argv=None):
    """Upstream hard-codes fold = 3 (C/train_CASIA.py:721); --fold k (1-based) selects another."""
    return run_main(FixedCASIACrossDomainTrainer, default_fold=3, argv=argv)


if __name__ == "__main__":
    main()
