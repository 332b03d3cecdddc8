"""CPU: the data layer reproduces the reference's on-disk format, fold splits and collator layout."""
import os
import sys

import numpy as np
import pytest
import torch

import dad_b200  # noqa: F401
from dad_b200 import dataload

LD = {"ang": 0, "hap": 1, "neu": 2, "sad": 3}


def _loop_collate(fs, idx):
    """The reference collator, restated sample by sample (I/dataload_clean.py:177-193)."""
    feats = [torch.from_numpy(np.array(fs.feats[fs.offsets[i]:fs.offsets[i] + fs.sizes[i]])).float() for i in idx]
    T = max(f.shape[0] for f in feats)
    out = feats[0].new_zeros(len(feats), T, 768)
    pad = torch.BoolTensor(len(feats), T).fill_(False)
    for i, f in enumerate(feats):
        out[i, :f.shape[0]] = f
        pad[i, f.shape[0]:] = True
    return out, pad


def test_vectorised_collator_equals_reference_layout():
    fs, _ = dataload.synthetic_feature_sets(5, 6, t_min=3, t_max=17, seed=3)
    tr, va, te, names, n = dataload.get_cv_dataloaders(None, 4, fold_id=1, label_dict=LD, feature_set=fs)
    assert names == list(LD) and n == 4
    seen = 0
    for b in va:  # unshuffled: positions map back to utterances of session 4 in file order
        ids = b["id"]
        src = np.where(fs.groups == 4)[0][ids.numpy()]
        feats, pad = _loop_collate(fs, src)
        assert torch.equal(b["net_input"]["feats"], feats) and torch.equal(b["net_input"]["padding_mask"], pad)
        assert torch.equal(b["labels"], torch.from_numpy(fs.labels[src]))
        assert b["net_input"]["feats"].dtype == torch.float32 and b["net_input"]["padding_mask"].dtype == torch.bool
        seen += len(ids)
    assert seen == 6 and len(va) == 2 and len(va.dataset) == 6  # ragged last batch kept, like DataLoader(drop_last=False)


def test_bucketed_padding_is_opt_in_and_only_appends_padding():
    """cfg.DAD_T_BUCKET / UtteranceLoader.t_bucket: default 0 = the reference's per-batch padding; b > 0 rounds the padded length
    up to a multiple of b and changes nothing but the number of trailing padded (zero, masked) frames."""
    fs, _ = dataload.synthetic_feature_sets(5, 6, t_min=3, t_max=17, seed=3)
    _, va, _, _, _ = dataload.get_cv_dataloaders(None, 4, fold_id=1, label_dict=LD, feature_set=fs)
    assert va.t_bucket == 0
    ref = [b for b in va]
    va.t_bucket = 8
    for b, r in zip(va, ref):
        x, pad, T0 = b["net_input"]["feats"], b["net_input"]["padding_mask"], r["net_input"]["feats"].shape[1]
        assert x.shape[1] % 8 == 0 and T0 <= x.shape[1] < T0 + 8
        assert torch.equal(x[:, :T0], r["net_input"]["feats"]) and torch.equal(pad[:, :T0], r["net_input"]["padding_mask"])
        assert bool(pad[:, T0:].all()) and float(x[:, T0:].abs().sum()) == 0.0


def test_fold_tables_and_speaker_isolation():
    assert dataload.iemocap_fold_sessions(1) == ([1, 2, 3], 4, 5) and dataload.iemocap_fold_sessions(5) == ([5, 1, 2], 3, 4)
    with pytest.raises(ValueError):
        dataload.iemocap_fold_sessions(0)
    names = [f"emo_{s}" for s in ("03", "08", "09", "10", "11", "12", "13", "14", "15", "16")]
    fs_c, fs_n = dataload.synthetic_feature_sets(10, 3, t_min=3, t_max=6, group_names=names)
    ids, train, val, test = dataload.speaker_fold(fs_c.groups, 3, emodb=True)
    assert (test, val) == ("10", "11") and len(train) == 8 and not ({val, test} & set(train))
    s, t, v, te = dataload.create_speaker_isolated_loaders(None, 4, 3, LD, noisy=True, emodb=True, feature_set=fs_n)
    b = next(iter(s))
    assert "labels" not in b and "id" not in b  # C/dataload_casia_noisy.py:72-107: SSL train batches carry neither
    assert next(iter(v))["labels"] is not None
    tr, va, te2, ncls, cn = dataload.create_speaker_isolated_loaders(None, 4, 3, LD, noisy=False, emodb=True, feature_set=fs_c)
    assert ncls == 4 and len(tr.dataset) == 24 and len(va.dataset) == 3 and len(te2.dataset) == 3


def test_on_disk_formats_roundtrip(tmp_path):
    fs, _ = dataload.synthetic_feature_sets(5, 2, t_min=3, t_max=5, seed=1)
    d = tmp_path / "iemocap"
    d.mkdir()
    np.save(d / "train.npy", fs.feats)
    (d / "train.lengths").write_text("".join(f"{s}\n" for s in fs.sizes))
    inv = {v: k for k, v in LD.items()}
    (d / "train.emo").write_text("".join(f"Ses0{g}_impro01_F{i:03d}\t{inv[int(y)]}\n" for i, (g, y) in enumerate(zip(fs.groups, fs.labels))))
    got = dataload.load_iemocap_set(str(d), LD)
    assert np.array_equal(got.sizes, fs.sizes) and np.array_equal(got.labels, fs.labels) and np.array_equal(got.groups.astype(int), fs.groups)
    assert np.array_equal(np.asarray(got.feats), fs.feats)
    tr, va, te, _, _ = dataload.get_cv_dataloaders(str(d), 3, fold_id=2, label_dict=LD)
    assert len(tr.dataset) == 6 and len(va.dataset) == 2 and len(te.dataset) == 2
    p = tmp_path / "casia"
    p.mkdir()
    np.save(p / "train.npy", fs.feats)
    (p / "train.lengths").write_text("".join(f"{s}\n" for s in fs.sizes))
    cl = {"angry": 0, "happy": 1, "neutral": 2, "sad": 3}
    invc = {v: k for k, v in cl.items()}
    (p / "train.lbl").write_text("".join(f"{invc[int(y)]}\n" for y in fs.labels))
    (p / "train.spk").write_text("".join(f"spk{g % 4}\n" for g in fs.groups))
    got = dataload.load_speaker_set(str(p / "train"), cl)
    assert np.array_equal(got.labels, fs.labels) and len(np.unique(got.groups)) == 4


@pytest.mark.skipif(not os.path.isdir("/root/reference/IEMOCAP/DAD-train-IEMOCAP"), reason="reference tree not present")
def test_collator_matches_live_reference():
    ref_dir = "/root/reference/IEMOCAP/DAD-train-IEMOCAP"
    sys.path.insert(0, ref_dir)
    sys.dont_write_bytecode = True
    try:
        import importlib
        ref = importlib.import_module("dataload_clean")
        fs, _ = dataload.synthetic_feature_sets(5, 4, t_min=3, t_max=9, seed=7)
        ds = ref.CleanEmotionDatasetFromArrays(fs.feats, fs.sizes, fs.offsets, fs.labels.tolist())
        idx = [5, 0, 13, 2]
        want = ds.collator([ds[i] for i in idx])
        ld = dataload.UtteranceLoader(fs, np.arange(len(fs)), 4, shuffle=False)
        got = ld.collate(torch.tensor(idx))
        assert torch.equal(got["net_input"]["feats"], want["net_input"]["feats"])
        assert torch.equal(got["net_input"]["padding_mask"], want["net_input"]["padding_mask"])
        assert torch.equal(got["labels"], want["labels"]) and torch.equal(got["id"], want["id"])
    finally:
        sys.path.remove(ref_dir)
        for m in ("dataload_clean", "config"):
            sys.modules.pop(m, None)
