"""DAD data layer: the reference's on-disk feature format, fold splits and batch layout.

Files (written by the reference's offline emotion2vec dump, pretrain-and-processed-*/scripts/emotion2vec_speech_features.py:94-119):
    <prefix>.npy       [sum T, 768] float32 frame features of all utterances, concatenated
    <prefix>.lengths   one integer per utterance
    <prefix>.emo       IEMOCAP: "<utt-name>\\t<label>" per utterance, session = int(name[4])    (I/dataload_clean.py:20-43,70-101)
    <prefix>.lbl/.spk  CASIA / EMODB: one label / one speaker id per utterance                   (C/dataload_casia_clean.py:19-68)

Splits: IEMOCAP 5 folds by session -- train 3 sessions, val 1, test 1 (I/dataload_clean.py:60-68); CASIA 4 folds / EMODB 10
folds by speaker -- test = fold-th speaker, val = the next one (C/dataload_casia_clean.py:70-110, E/dataload_emodb_noisy.py:24-49).

Batches are the reference collator's dicts (I/dataload_clean.py:177-193): zero-padded to the batch's longest utterance,
`padding_mask` True = pad; noisy TRAIN batches carry no labels (SSL; C/dataload_casia_noisy.py:72-107 has no 'id' / 'labels').
The loaders here keep the whole split resident as ONE pinned host array and assemble a batch with a single vectorised
gather (the reference loops over samples in Python); `num_workers=0`, shuffle per epoch, last batch ragged -- as upstream.
`synthetic_feature_sets` provides class-structured stand-in data of the same layout for offline runs and tests."""
import os

import numpy as np
import torch

D_IN = 768


class FeatureSet:
    """One domain's utterances: feats [sum T, 768] f32, sizes / offsets [U], labels [U] (int64, -1 = unknown), groups [U]."""

    def __init__(self, feats, sizes, labels, groups):
        self.feats = feats
        self.sizes = np.asarray(sizes, dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)[:-1]]).astype(np.int64)
        self.labels = np.asarray(labels, dtype=np.int64)
        self.groups = np.asarray(groups)

    def __len__(self):
        return len(self.sizes)


def load_iemocap_set(data_dir, label_dict):
    """<data_dir>/train.{npy,lengths,emo}; session id = 5th character of the utterance name (I/dataload_clean.py:35-37)."""
    prefix = os.path.join(data_dir, "train")
    feats = np.load(prefix + ".npy", mmap_mode="r")
    with open(prefix + ".lengths") as f:
        sizes = [int(line) for line in f if line.strip()]
    labels, sessions = [], []
    with open(prefix + ".emo", encoding="utf-8") as f:
        for line in f:
            if not line.strip():
                continue
            parts = line.rstrip().split()
            labels.append(label_dict[parts[1]])
            sessions.append(int(parts[0][4]))
    if not (len(sizes) == len(labels)):
        raise ValueError(f"{prefix}: {len(sizes)} lengths vs {len(labels)} labels")
    return FeatureSet(feats, sizes, labels, sessions)


def load_speaker_set(prefix, label_dict):
    """<prefix>.{npy,lengths,lbl,spk} (CASIA / EMODB)."""
    feats = np.load(prefix + ".npy", mmap_mode="r")
    with open(prefix + ".lengths") as f:
        sizes = [int(line) for line in f if line.strip()]
    with open(prefix + ".lbl", encoding="utf-8") as f:
        labels = [label_dict[line.strip()] for line in f if line.strip()]
    with open(prefix + ".spk", encoding="utf-8") as f:
        speakers = [line.strip() for line in f if line.strip()]
    if not (len(sizes) == len(labels) == len(speakers)):
        raise ValueError(f"{prefix}: inconsistent sidecar lengths")
    return FeatureSet(feats, sizes, labels, speakers)


def iemocap_fold_sessions(fold_id):
    """fold_id 1..5 -> (train sessions, val session, test session).  I/dataload_clean.py:60-68."""
    if not 1 <= fold_id <= 5:
        raise ValueError(f"fold_id must be between 1 and 5, got {fold_id}")
    table = {1: ([1, 2, 3], 4, 5), 2: ([2, 3, 4], 5, 1), 3: ([3, 4, 5], 1, 2), 4: ([4, 5, 1], 2, 3), 5: ([5, 1, 2], 3, 4)}
    return table[fold_id]


def speaker_fold(groups, fold, emodb=False):
    """fold (0-based) -> (train ids, val id, test id) over the sorted unique speakers; EMODB compares the id after the last '_'."""
    ids = np.array([g.split("_")[-1] for g in groups]) if emodb else np.asarray(groups)
    uniq = np.unique(ids)
    if not 0 <= fold < len(uniq):
        raise ValueError(f"fold must be in [0, {len(uniq)}), got {fold}")
    test, val = uniq[fold], uniq[(fold + 1) % len(uniq)]
    train = [u for u in uniq if u not in (test, val)]
    return ids, train, val, test


class UtteranceLoader:
    """Iterable over padded batches of one split, with the reference DataLoader's observable behaviour
    (`len()` = number of batches incl. the ragged last one, `.dataset` with `len()`, reshuffle on every `iter()`)."""

    class _Dataset:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

    def __init__(self, fs, indices, batch_size, shuffle, use_labels=True, with_id=True, generator=None):
        idx = np.asarray(indices, dtype=np.int64)
        sizes = fs.sizes[idx]
        # one contiguous, pinned host copy of the split (the reference concatenates per split too, I/dataload_clean.py:255-262)
        rows = np.concatenate([np.arange(fs.offsets[i], fs.offsets[i] + fs.sizes[i]) for i in idx]) if len(idx) else np.zeros(0, np.int64)
        feats = torch.from_numpy(np.ascontiguousarray(fs.feats[rows])).float() if len(idx) else torch.zeros(0, D_IN)
        self.feats = feats.pin_memory() if torch.cuda.is_available() else feats
        self.sizes = torch.from_numpy(sizes)
        self.offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)) if len(idx) else torch.zeros(0, dtype=torch.long)
        self.labels = torch.from_numpy(fs.labels[idx])
        self.batch_size, self.shuffle, self.use_labels, self.with_id = int(batch_size), shuffle, use_labels, with_id
        self.generator = generator
        self.dataset = self._Dataset(len(idx))

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    # 0 = pad every batch to ITS OWN longest utterance, exactly as the reference collator (I/dataload_clean.py:177-193).  b > 0 pads to
    # the next multiple of b frames (cfg.DAD_T_BUCKET, training loaders only): shuffled batches then fall into a few dozen padded
    # lengths and the step engine replays captured CUDA graphs instead of enqueueing ~30 launches eagerly per step.  Opt-in because
    # the strong augmentation's span length is int(T_padded * TEMPORAL_MASK_RATIO) (I/utils.py:366): bucketing moves it by < b/10 frames.
    t_bucket = 0

    def _padded_len(self, t_max):
        b = int(self.t_bucket)
        return t_max if b <= 0 else (t_max + b - 1) // b * b

    def collate(self, ids):
        """ids [b] (positions inside the split) -> reference batch dict (I/dataload_clean.py:177-193), one vectorised gather."""
        sizes, offs = self.sizes[ids], self.offsets[ids]
        b, T = len(ids), self._padded_len(int(sizes.max()))
        t = torch.arange(T)
        pad = t[None, :] >= sizes[:, None]
        src = (offs[:, None] + t[None, :]).masked_fill(pad, 0)
        feats = self.feats[src.reshape(-1)].view(b, T, D_IN)
        feats.masked_fill_(pad.unsqueeze(-1), 0.0)
        batch = {"net_input": {"feats": feats, "padding_mask": pad}}
        if self.with_id:
            batch["id"] = ids.clone()
        if self.use_labels:
            batch["labels"] = self.labels[ids].clone()
        elif self.with_id:
            batch["labels"] = None
        return batch

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        for s in range(0, n, self.batch_size):
            yield self.collate(order[s:s + self.batch_size])


class DeviceUtteranceLoader(UtteranceLoader):
    """Same interface, but the split's features live in HBM and every batch is assembled ON the device by one gather kernel
    (csrc/k_aug.cu k_gather_pad): no per-sample Python loop, no pageable host->device copy per step
    (I/dataload_clean.py:170-193, I/train.py:398,405).  An IEMOCAP training fold is ~2.2 GB of the 180 GB."""

    def __init__(self, fs, indices, batch_size, shuffle, use_labels=True, with_id=True, generator=None, device=None, sink=None):
        super().__init__(fs, indices, batch_size, shuffle, use_labels, with_id, generator)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.feats_dev = self.feats.to(self.device, non_blocking=True)
        self.offsets_dev = self.offsets.to(self.device)
        self.sizes_dev = self.sizes.to(torch.int32).to(self.device)
        self.labels_dev = self.labels.to(self.device)
        # sink(B, T) -> (x [B,T,768] f32, pad [B,T] u8): where the gathered batch is WRITTEN.  The trainers point it at the step
        # engine's own input buffers (DADStepEngine.input_views), so a batch is assembled in place and set_inputs copies nothing.
        self.sink = sink

    @classmethod
    def from_device_batch(cls, batch, batch_size, sink=None, use_labels=True):
        """A resident split made of the utterances of one padded device batch [B,T,768] (benchmarks / tests): every utterance keeps
        its valid length; the feature rows are a view of the batch (no copy)."""
        self = cls.__new__(cls)
        x = batch["net_input"]["feats"]
        B, T = int(x.shape[0]), int(x.shape[1])
        pm = batch["net_input"].get("padding_mask")
        sizes = torch.full((B,), T, dtype=torch.int64) if pm is None else (~pm.bool()).sum(1).cpu().to(torch.int64)
        self.device = x.device
        self.feats = self.feats_dev = x.reshape(B * T, D_IN)
        self.sizes = sizes
        self.offsets = torch.arange(B, dtype=torch.int64) * T
        lab = batch.get("labels")
        self.labels = lab.cpu() if lab is not None else torch.zeros(B, dtype=torch.int64)
        self.offsets_dev, self.sizes_dev = self.offsets.to(x.device), sizes.to(torch.int32).to(x.device)
        self.labels_dev = self.labels.to(x.device)
        self.batch_size, self.shuffle, self.use_labels, self.with_id, self.generator = int(batch_size), False, use_labels and lab is not None, True, None
        self.dataset = self._Dataset(B)
        self.sink = sink
        return self

    def collate(self, ids):
        ids_dev = ids.to(self.device, non_blocking=True)
        return self._assemble(ids, self.offsets_dev[ids_dev], self.sizes_dev[ids_dev], self.labels_dev[ids_dev] if self.use_labels else None,
                              int(self.sizes[ids].max()))  # host copy of the lengths: no device sync

    def _assemble(self, ids, offs_dev, sizes_dev, labels_dev, t_max):
        from . import ops
        T = self._padded_len(t_max)
        x_out = pad_out = None
        if self.sink is not None:
            x_out, pad_out = self.sink(len(ids), T)
        feats, pad = ops.gather_pad(self.feats_dev, offs_dev, sizes_dev, T=T, x_out=x_out, pad_out=pad_out)
        batch = {"net_input": {"feats": feats, "padding_mask": pad}}
        if self.with_id:
            batch["id"] = ids.clone()
        if self.use_labels:
            batch["labels"] = labels_dev
        elif self.with_id:
            batch["labels"] = None
        return batch

    def __iter__(self):
        """Same order and batches as UtteranceLoader.__iter__; the epoch's permutation is uploaded ONCE and the per-utterance
        offsets / lengths / labels are gathered once per epoch, so a batch costs slicing + one gather kernel on the host side (the
        per-batch index uploads and gathers were ~0.16 ms of host time per batch -- more than the step's GPU time at batch 64)."""
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        order_dev = order.to(self.device, non_blocking=True)
        offs, sizes = self.offsets_dev[order_dev], self.sizes_dev[order_dev]
        labels = self.labels_dev[order_dev] if self.use_labels else None
        bs = self.batch_size
        sizes_host = self.sizes[order]
        t_max = [int(v) for v in torch.nn.functional.pad(sizes_host, (0, (-n) % bs)).view(-1, bs).max(dim=1).values] if n else []
        for i, s in enumerate(range(0, n, bs)):
            yield self._assemble(order[s:s + bs], offs[s:s + bs], sizes[s:s + bs], labels[s:s + bs] if labels is not None else None, t_max[i])


def _split_loaders(fs, train_idx, val_idx, test_idx, batch_size, noisy, with_id, device=None):
    cls = UtteranceLoader if device is None else (lambda *a, **k: DeviceUtteranceLoader(*a, device=device, **k))
    mk = lambda idx, shuffle, labels: cls(fs, idx, batch_size, shuffle, use_labels=labels, with_id=with_id)
    if noisy:  # student + teacher train loaders are two independently shuffled views of the same unlabeled split
        return mk(train_idx, True, False), mk(train_idx, True, False), mk(val_idx, False, True), mk(test_idx, False, True)
    return mk(train_idx, True, True), mk(val_idx, False, True), mk(test_idx, False, True)


# ------------------------------------------------------------------------------------------------ reference entry points
def get_cv_dataloaders(data_path, batch_size, fold_id=1, label_dict=None, feature_set=None, device=None):
    """IEMOCAP clean: -> (train_loader, val_loader, test_loader, class_names, num_classes).  I/dataload_clean.py:222-295."""
    from . import config as cfg
    label_dict = label_dict or cfg.LABEL_DICT
    fs = feature_set if feature_set is not None else load_iemocap_set(data_path, label_dict)
    tr, va, te = iemocap_fold_sessions(fold_id)
    g = fs.groups.astype(np.int64)
    loaders = _split_loaders(fs, np.where(np.isin(g, tr))[0], np.where(g == va)[0], np.where(g == te)[0], batch_size, False, True, device)
    return (*loaders, list(label_dict.keys()), len(label_dict))


def get_cv_dataloaders_noisy(data_path, batch_size, fold_id=1, label_dict=None, feature_set=None, device=None):
    """IEMOCAP noisy: -> (student_loader, teacher_loader, val_loader, test_loader).  I/dataload_noisy.py:159-245."""
    from . import config as cfg
    label_dict = label_dict or cfg.LABEL_DICT
    fs = feature_set if feature_set is not None else load_iemocap_set(data_path, label_dict)
    tr, va, te = iemocap_fold_sessions(fold_id)
    g = fs.groups.astype(np.int64)
    return _split_loaders(fs, np.where(np.isin(g, tr))[0], np.where(g == va)[0], np.where(g == te)[0], batch_size, True, True, device)


def create_speaker_isolated_loaders(feat_path, batch_size, fold, label_dict, noisy, emodb=False, feature_set=None, device=None):
    """CASIA (4 folds) / EMODB (10 folds): clean -> (train, val, test, num_classes, class_names); noisy -> (student, teacher,
    val, test).  C/dataload_casia_clean.py:70-160, C/dataload_casia_noisy.py:160-315, E/dataload_emodb_noisy.py:191-376."""
    fs = feature_set if feature_set is not None else load_speaker_set(feat_path, label_dict)
    ids, train, val, test = speaker_fold(fs.groups, fold, emodb=emodb)
    tr_idx, va_idx, te_idx = np.where(np.isin(ids, train))[0], np.where(ids == val)[0], np.where(ids == test)[0]
    assert not (set(train) & {val, test}) and val != test, "speaker leak between splits"  # E/dataload_emodb_noisy.py:253-255
    loaders = _split_loaders(fs, tr_idx, va_idx, te_idx, batch_size, noisy, with_id=not noisy, device=device)
    if noisy:
        return loaders
    return (*loaders, len(label_dict), list(label_dict.keys()))


# ------------------------------------------------------------------------------------------------ synthetic stand-in data
def synthetic_feature_sets(n_groups, utt_per_group, t_min=40, t_max=120, snr_db=10.0, seed=0, group_names=None):
    """-> (clean FeatureSet, noisy FeatureSet) with the synthetic distribution of dad_b200.synth: class means
    0.5*randn(4,768) + unit noise; the noisy domain adds white noise at `snr_db` per utterance.  Groups = sessions (1..n) or
    the given speaker names; both domains hold the same utterance list (as the reference's parallel clean / noisy dumps)."""
    from . import synth
    g = torch.Generator().manual_seed(seed)
    mu = synth.class_means()
    U = n_groups * utt_per_group
    sizes = torch.randint(t_min, t_max + 1, (U,), generator=g)
    labels = torch.randint(0, 4, (U,), generator=g)
    groups = np.repeat(np.arange(1, n_groups + 1) if group_names is None else np.asarray(group_names), utt_per_group)
    clean_rows, noisy_rows = [], []
    for u in range(U):
        T = int(sizes[u])
        x = mu[labels[u]][None, :] + torch.randn(T, D_IN, generator=g)
        x2 = mu[labels[u]][None, :] + torch.randn(T, D_IN, generator=g)
        sigma = torch.sqrt((x2.double() ** 2).mean() / (10.0 ** (snr_db / 10.0))).float()
        clean_rows.append(x)
        noisy_rows.append(x2 + sigma * torch.randn(T, D_IN, generator=g))
    mk = lambda rows: FeatureSet(torch.cat(rows).numpy(), sizes.numpy(), labels.numpy(), groups)
    return mk(clean_rows), mk(noisy_rows)
