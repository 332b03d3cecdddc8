"""Trainer-level throughput at the reference's own configuration (batch 64, shuffled variable-length utterances, per-batch padding):
the loop a user of `train.py --fold k` runs -- DeviceUtteranceLoader.collate -> engine.set_inputs -> engine.step per batch, one loss
read-back per epoch -- timed by wall clock around whole epochs with a device sync at both ends.
Reports samples/s after the warm-up phase, the engine's graph statistics (eager / captured / replayed steps) and, with --bucket B, the
same run with training batches padded to multiples of B frames (cfg.DAD_T_BUCKET) so that step shapes repeat.
Usage: python tools/trainer_bench.py [--utt-per-session 1024] [--t-min 100] [--t-max 600] [--epochs 3] [--bucket 32]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dad_b200  # noqa: F401
from dad_b200 import config as cfg, dataload
from dad_b200.train import IEMOCAPCrossDomainTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--utt-per-session", type=int, default=1024)
ap.add_argument("--t-min", type=int, default=100)
ap.add_argument("--t-max", type=int, default=600)
ap.add_argument("--epochs", type=int, default=3, help="timed post-warm-up epochs")
ap.add_argument("--bucket", type=int, default=0)
ap.add_argument("--batch", type=int, default=64)
args = ap.parse_args()

cfg.BATCH_SIZE = args.batch
cfg.DAD_T_BUCKET = args.bucket
cfg.ANCHOR_CALIBRATION_ENABLED = False
cfg.EPOCHS = max(int(cfg.EPOCHS), int(cfg.WARMUP_EPOCHS) + args.epochs + 1)
dev = torch.device("cuda", 0)
fs_c, fs_n = dataload.synthetic_feature_sets(5, args.utt_per_session, t_min=args.t_min, t_max=args.t_max, seed=0)


def loaders(batch_size):
    return (dataload.get_cv_dataloaders(None, batch_size, 1, cfg.LABEL_DICT, feature_set=fs_c, device=dev),
            dataload.get_cv_dataloaders_noisy(None, batch_size, 1, cfg.LABEL_DICT, feature_set=fs_n, device=dev))


tr = IEMOCAPCrossDomainTrainer(fold=0, experiment_name="trainer_bench", loaders=loaders, require_pretrained=False)
eng = tr.engine
n_batches = min(len(tr.clean_train_loader), len(tr.noisy_student_loader))
e_full = int(cfg.WARMUP_EPOCHS) + int(getattr(cfg, "WEIGHT_RAMP_EPOCHS", 10)) + 5  # an epoch with the full loss stack active
tr.train_epoch(0)            # CE-only warm-up epoch (untimed)
tr.train_epoch(e_full)       # first full-stack epoch (untimed: cold graph cache)
torch.cuda.synchronize()
stats0 = dict(eng.graph_stats)
t0 = time.perf_counter()
for k in range(args.epochs):
    tr.train_epoch(e_full)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
stats = {k: eng.graph_stats[k] - stats0.get(k, 0) for k in eng.graph_stats}
steps = n_batches * args.epochs
print(json.dumps({"what": "trainer-level DAD training throughput (full loss stack, HBM-resident split, per-batch padding)",
                  "batch": args.batch, "t_range": [args.t_min, args.t_max], "t_bucket": args.bucket, "steps": steps,
                  "samples_per_s": round(steps * args.batch / dt, 1), "ms_per_step": round(dt / steps * 1e3, 4),
                  "graph_stats": stats, "distinct_shapes_seen": len(eng._key_seen) + len(eng._graphs)}))
