"""CPU: the oracle's SupCon restatement (parity UNPINNED -- the reference removed its SCL, C/utils.py:377-378) against a
literal fp64 loop over the published formula (Khosla et al. 2020, eq. 2: L_out with mean over positives)."""
import math

import torch

from oracle import dad_oracle as O


def _literal(feats, labels, member, tau):
    idx = [i for i in range(len(labels)) if member[i]]
    f = [feats[i].double() / feats[i].double().norm() for i in idx]
    y = [int(labels[i]) for i in idx]
    tot, n = 0.0, 0
    for a in range(len(idx)):
        pos = [p for p in range(len(idx)) if p != a and y[p] == y[a]]
        if not pos:
            continue
        den = sum(math.exp(float(f[a] @ f[b]) / tau) for b in range(len(idx)) if b != a)
        tot += -sum(float(f[a] @ f[p]) / tau - math.log(den) for p in pos) / len(pos)
        n += 1
    return tot / n if n else 0.0


def test_scl_restatement_matches_literal_formula():
    g = torch.Generator().manual_seed(4)
    feats = torch.randn(20, 256, generator=g, dtype=torch.float64)
    labels = torch.randint(0, 4, (20,), generator=g)
    member = torch.rand(20, generator=g) > 0.25
    assert abs(float(O.scl_loss(feats, labels, member, 0.1)) - _literal(feats, labels, member, 0.1)) < 1e-9
    lone = torch.tensor([0, 1, 2, 3])  # no anchor has a positive -> 0
    assert float(O.scl_loss(feats[:4], lone, torch.ones(4, dtype=torch.bool), 0.1)) == 0.0
