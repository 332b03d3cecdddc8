"""Shared test plumbing: golden fixtures, oracle construction, engine<->oracle state transfer."""
import os
import types

import torch

import dad_b200  # noqa: F401
from dad_b200 import synth
from oracle import dad_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STUDENT = O.PARAM_KEYS


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def pretrained_state():
    s = load_golden("pretrained_student")
    out = dict(s)
    out.update({k.replace("student_", "teacher_"): v.clone() for k, v in s.items()})
    return out


def cfg_for(dataset, overrides=None):
    """A private copy of the dataset's config module namespace (so tests can override fields freely)."""
    from dad_b200 import config, config_casia, config_emodb
    mod = {"iemocap": config, "casia": config_casia, "emodb": config_emodb}[dataset]
    ns = types.SimpleNamespace(**{k: v for k, v in vars(mod).items() if k.isupper()})
    for k, v in (overrides or {}).items():
        setattr(ns, k, v)
    return ns


def make_oracle(dataset, overrides=None, state=None):
    cfg = cfg_for(dataset, overrides)
    hp = O.HParams.from_cfg(cfg)
    return O.OracleStepper(state if state is not None else pretrained_state(), hp, dataset=dataset), cfg, hp


def step_inputs(meta, k, cfg):
    c, n = synth.make_batches(meta["B"], meta["T"], meta["seed"] + k, snr_db=meta["snr_db"], variable_len=meta["variable_len"])
    rng = synth.make_rng_inputs(meta["B"], meta["T"], meta["seed"] + 100000 + k, drop_p=cfg.DROPOUT_RATE,
                                mask_ratio=cfg.TEMPORAL_MASK_RATIO)
    return c, n, rng


def flat_from(d, keys):
    return torch.cat([d[k].detach().reshape(-1).float() for k in keys])


def push_oracle_state(engine, orc):
    """Teacher forcing: copy the oracle's full optimisation state into the engine."""
    engine.p.copy_(flat_from(orc.p, O.PARAM_KEYS))
    engine.t.copy_(flat_from(orc.t, O.TEACHER_KEYS))
    engine.m.copy_(flat_from(orc.m, O.PARAM_KEYS))
    engine.v.copy_(flat_from(orc.v, O.PARAM_KEYS))
    engine.step_count.fill_(orc.nstep)
    engine.ema_thr.copy_(orc.dacp.ema_thr)
    engine.quality.copy_(orc.dacp.quality)
    engine.anchors.copy_(orc.anchors)
    engine.set_lr(orc.lr)
    engine.refresh_splits()


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)
