"""torchrun worker for the data-parallel tests: batch-sharded step == single-process global-batch step.

    torchrun --nproc-per-node W tests/dp_worker.py --mode cpu    # gloo, host logic: sharded ORACLE protocol vs the oracle
    torchrun --nproc-per-node W tests/dp_worker.py --mode gpu    # nccl, the CUDA engine world=W vs world=1 on each rank

cpu mode checks the PROTOCOL the engine implements -- per-rank forwards, ONE packed all-gather in the dad_pack layout
(dad_b200._lib.pack_offsets / pack_global_views), DACP + ECDA redundantly on the global batch, gradients only through
the local slice, ONE all-reduce(sum) of the flat gradient -- with oracle arithmetic, against the reference-pinned
single-process oracle step on the global batch.  gpu mode runs the product."""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dad_b200  # noqa: E402,F401
from dad_b200 import _lib, synth  # noqa: E402
from helpers import cfg_for, pretrained_state  # noqa: E402
from oracle import dad_oracle as O  # noqa: E402


def shard(batch, r, B):
    sl = slice(r * B, (r + 1) * B)
    return {"net_input": {"feats": batch["net_input"]["feats"][sl].contiguous(),
                          "padding_mask": batch["net_input"]["padding_mask"][sl].contiguous()}, "labels": batch["labels"][sl].contiguous()}


def shard_rng(rng, r, B):
    sl = slice(r * B, (r + 1) * B)
    out = {k: rng[k][sl].contiguous() for k in ("drop_clean", "n_weak", "n_strong", "t_start", "drop_strong")}
    out["feat_keep_u"] = rng["feat_keep_u"]  # one [768] draw shared by the whole (global) batch, I/utils.py:343
    return out


# ------------------------------------------------------------------------------------------------ cpu / gloo
def sharded_oracle_step(orc, c, n, epoch, rng, rank, world):
    """One DP step of rank `rank` with oracle arithmetic; mutates `orc` (replicated state) exactly like OracleStepper.step."""
    hp, dt = orc.hp, orc.dtype
    B = c["labels"].numel()
    Bg = B * world
    W, b, Wc, bc = (orc.p[k] for k in O.PARAM_KEYS)
    tW, tb, tWc, tbc = (orc.t[k] for k in O.TEACHER_KEYS)
    for p in orc.p.values():
        p.grad = None
    eps = hp.LABEL_SMOOTHING_FACTOR if hp.USE_LABEL_SMOOTHING else 0.0
    emb_c = O.encoder(c["net_input"]["feats"], c["net_input"]["padding_mask"], W, b)
    logits_c = O.classifier(emb_c, rng["drop_clean"], hp.DROPOUT_RATE, Wc, bc)
    ce_share = F.cross_entropy(logits_c, c["labels"], label_smoothing=eps, reduction="sum") / Bg
    post = epoch >= hp.WARMUP_EPOCHS
    obj = ce_share
    shares = torch.zeros(3, dtype=dt)
    shares[0] = ce_share.detach()
    if post:
        w_kd, w_ecda = O.loss_weights(epoch, hp)
        xn, pn = n["net_input"]["feats"], n["net_input"]["padding_mask"]
        weak = O.weak_augment(xn, rng["n_weak"], hp.WEAK_NOISE_STD)
        strong = O.strong_augment(xn, rng["n_strong"], hp.STRONG_NOISE_STD, rng["feat_keep_u"], hp.DROPOUT_RATE, rng["t_start"],
                                  hp.TEMPORAL_MASK_RATIO)
        with torch.no_grad():
            probs = F.softmax(O.classifier(O.encoder(weak, pn, tW, tb), None, 0.0, tWc, tbc), dim=1)
            score, pred = O.certainty_scores(probs, hp.USE_ENTROPY_IN_SCORE)
        emb_s = O.encoder(strong, pn, W, b)
        logits_s = O.classifier(emb_s, rng["drop_strong"], hp.DROPOUT_RATE, Wc, bc)
        # ---- the ONE packed all-gather (dad_pack layout)
        words = _lib.pack_rank_words(B)
        pack = torch.zeros(world * words, dtype=torch.float32)
        gv = _lib.pack_global_views(pack, B, world)
        gv["emb_clean"][rank].copy_(emb_c.detach()); gv["emb_strong"][rank].copy_(emb_s.detach())
        gv["score"][rank].copy_(score); gv["pred"][rank].copy_(pred.int()); gv["label"][rank].copy_(c["labels"].int())
        dist.all_gather_into_tensor(pack, pack[rank * words:(rank + 1) * words].clone())
        g_score, g_pred = gv["score"].reshape(Bg), gv["pred"].reshape(Bg).long()
        g_label = gv["label"].reshape(Bg).long()
        # ---- DACP redundantly on the global batch (I/utils.py:449-507, restated on gathered scores / pseudo-labels)
        st = orc.dacp
        wq = torch.sigmoid(hp.DACP_SENSITIVITY_K * (st.quality - st.quality.mean()))
        gamma = hp.DACP_QUANTILE_START + (hp.DACP_QUANTILE_END - hp.DACP_QUANTILE_START) * (epoch / hp.EPOCHS)
        tau = torch.stack([torch.quantile(g_score[g_pred == k], gamma) if (g_pred == k).any() else st.ema_thr[k] for k in range(4)])
        tau = torch.max(tau + hp.DACP_CALIBRATION_STRENGTH_LAMBDA * (wq - 0.5), orc.anchors)
        a = hp.DACP_THRESHOLD_SMOOTHING_ALPHA
        st.ema_thr = a * st.ema_thr + (1 - a) * tau
        g_mask = g_score >= st.ema_thr[g_pred]
        for k in range(4):
            st.epoch_scores[k].extend(g_score[g_pred == k].tolist())
        nmask = g_mask.sum()
        sl = slice(rank * B, (rank + 1) * B)
        if int(nmask) > 1:
            logq = F.log_softmax(logits_s, dim=1)
            kl = F.kl_div(logq, probs, reduction="none").sum(dim=1)
            kd_share = (kl * g_mask[sl].to(dt)).sum() / (nmask.to(dt) + 1e-8)
            shares[1] = kd_share.detach()
            obj = obj + w_kd * kd_share
            if hp.USE_ECDA and w_ecda > 0:
                # global embeddings with the local slice attached to the autograd graph: d/d(local rows) only
                ge_c = torch.cat([gv["emb_clean"].reshape(Bg, -1)[:rank * B], emb_c, gv["emb_clean"].reshape(Bg, -1)[(rank + 1) * B:]])
                ge_s = torch.cat([gv["emb_strong"].reshape(Bg, -1)[:rank * B], emb_s, gv["emb_strong"].reshape(Bg, -1)[(rank + 1) * B:]])
                ecda = O.ecda_loss(ge_c, ge_s, g_label, g_pred, g_mask, g_score, wq, hp)
                shares[2] = ecda.detach() / world  # every rank holds the full value; shares sum to it
                obj = obj + w_ecda * ecda
        orc.last = dict(mask=g_mask, pred=g_pred)
    obj.backward()
    flat = torch.cat([orc.p[k].grad.reshape(-1) for k in O.PARAM_KEYS] + [shares])
    dist.all_reduce(flat)  # ---- the ONE all-reduce: gradient ++ loss shares
    off = 0
    for k in O.PARAM_KEYS:
        nel = orc.p[k].numel()
        orc.p[k].grad = flat[off:off + nel].view_as(orc.p[k]).clone()
        off += nel
    ce, kd, ecda = (float(v) for v in flat[off:off + 3])
    if hp.GRADIENT_CLIPPING:
        O.clip_grad_norm([orc.p[k].grad for k in O.PARAM_KEYS], hp.MAX_GRAD_NORM)
    orc.nstep += 1
    with torch.no_grad():
        for k in O.PARAM_KEYS:
            O.adam_step(orc.p[k], orc.p[k].grad, orc.m[k], orc.v[k], orc.nstep, orc.lr, hp.WEIGHT_DECAY)
        if post:
            for ks, kt in zip(O.PARAM_KEYS, O.TEACHER_KEYS):
                orc.t[kt] = O.ema_update(orc.t[kt], orc.p[ks].detach(), hp.EMA_MOMENTUM)
    w_kd, w_ecda = O.loss_weights(epoch, hp)
    return {"supervised_ce_loss": ce, "consistency_loss": kd, "ecda_loss": ecda, "total_loss": ce + w_kd * kd + w_ecda * ecda}


def run_cpu(args):
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.set_num_threads(2)
    B, T, steps = args.B, args.T, args.steps
    Bg = B * world
    cfg = cfg_for("iemocap")
    hp = O.HParams.from_cfg(cfg)
    glob = O.OracleStepper(pretrained_state(), hp)
    mine = O.OracleStepper(pretrained_state(), hp)
    worst = 0.0
    for k in range(steps):
        epoch = 5 if k == 0 else 40 + k  # one warm-up (CE-only) step, then the full loss stack
        glob.lr = mine.lr = O.cosine_lr(hp.LEARNING_RATE, epoch, hp.EPOCHS)
        c, n = synth.make_batches(Bg, T, 4000 + k, variable_len=True)
        rng = synth.make_rng_inputs(Bg, T, 5000 + k)
        ref = glob.step(c, n, epoch, rng)
        got = sharded_oracle_step(mine, shard(c, rank, B), shard(n, rank, B), epoch, shard_rng(rng, rank, B), rank, world)
        for key in ("total_loss", "supervised_ce_loss", "consistency_loss", "ecda_loss"):
            r = abs(got[key] - ref[key]) / max(abs(ref[key]), 1e-3)
            worst = max(worst, r)
            assert r <= 2e-5, (rank, k, key, got[key], ref[key])
        if epoch >= hp.WARMUP_EPOCHS:
            assert torch.equal(mine.last["mask"], glob.last["mask"]), (rank, k, "mask")
            assert torch.equal(mine.last["pred"], glob.last["pred"]), (rank, k, "pred")
            assert float((mine.dacp.ema_thr - glob.dacp.ema_thr).abs().max()) <= 1e-7
        for kk in O.PARAM_KEYS:
            g_ok = glob.p[kk].grad.abs() >= 1e-6  # ill-conditioned Adam elements excluded, as in tests/test_gpu_step.py
            d = (mine.p[kk].detach() - glob.p[kk].detach()).abs()
            assert float(d[g_ok].max() if g_ok.any() else 0.0) <= 1e-5, (rank, k, kk, float(d.max()))
    # replicated state must be IDENTICAL on every rank (SURVEY.md 8e "Determinism")
    flat = torch.cat([mine.p[k].detach().reshape(-1) for k in O.PARAM_KEYS] + [mine.dacp.ema_thr])
    ref0 = flat.clone()
    dist.broadcast(ref0, 0)
    assert torch.equal(flat, ref0), (rank, "replicated state diverged across ranks")
    if rank == 0:
        print(f"DP_CPU_OK world={world} steps={steps} worst loss rel={worst:.2e}")
    dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ gpu / nccl
def run_gpu(args):
    from dad_b200.engine import DADStepEngine
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    B, T, steps = args.B, args.T, args.steps
    Bg = B * world
    cfg = cfg_for("iemocap")
    gm = _lib.GEMM_TC if args.gemm == "tc" else _lib.GEMM_FP32
    dp = DADStepEngine(cfg, B, T, device=dev, world=world, rank=rank, process_group=dist.group.WORLD, gemm_mode=gm,
                       use_graph=args.graph, noise_mode="supplied", exchange=args.exchange)
    assert dp.exchange == args.exchange, (dp.exchange, args.exchange)
    one = DADStepEngine(cfg, Bg, T, device=dev, gemm_mode=gm, use_graph=False, noise_mode="supplied")
    st = pretrained_state()
    dp.load_state_dict(st); one.load_state_dict(st)
    cu = lambda b: {"net_input": {k: v.to(dev) for k, v in b["net_input"].items()}, "labels": b["labels"].to(dev)}
    worst = {"loss": 0.0, "w": 0.0}
    for k in range(steps):
        epoch = 5 if k == 0 else 40 + k
        lr = O.cosine_lr(cfg.LEARNING_RATE, epoch, cfg.EPOCHS)
        dp.set_lr(lr); one.set_lr(lr)
        c, n = synth.make_batches(Bg, T, 4000 + k, variable_len=True)
        rng = synth.make_rng_inputs(Bg, T, 5000 + k)
        one.set_inputs(cu(c), cu(n), {kk: v.to(dev) for kk, v in rng.items()})
        dp.set_inputs(cu(shard(c, rank, B)), cu(shard(n, rank, B)), {kk: v.to(dev) for kk, v in shard_rng(rng, rank, B).items()})
        one.step(epoch); dp.step(epoch)
        a, b = dp.losses(), one.losses()
        for key in a:
            r = abs(a[key] - b[key]) / max(abs(b[key]), 1e-3)
            worst["loss"] = max(worst["loss"], r)
            assert r <= 1e-4, (rank, k, key, a[key], b[key])
        if epoch >= cfg.WARMUP_EPOCHS:
            assert torch.equal(dp.mask, one.mask), (rank, k, "mask")
            assert float((dp.ema_thr - one.ema_thr).abs().max()) <= 1e-6
        well = one.g_ext[:_lib.NPARAM].abs() >= 1e-6
        d = (dp.p - one.p).abs()
        worst["w"] = max(worst["w"], float(d[well].max()))
        assert float(d[well].max()) <= 1e-5, (rank, k, "student weights", float(d[well].max()))
        assert float((dp.t - one.t).abs().max()) <= 1e-5, (rank, k, "teacher weights")
        # keep the comparison teacher-forced: DP engine state := single-process state
        for name in ("p", "t", "m", "v", "ema_thr"):
            getattr(dp, name).copy_(getattr(one, name))
        dp.refresh_splits()
    # replicated state bit-identical across ranks after free-running steps
    for k in range(3):
        dp.step(45)
    flat = torch.cat([dp.p, dp.t, dp.ema_thr])
    ref0 = flat.clone()
    dist.broadcast(ref0, 0)
    assert torch.equal(flat, ref0), (rank, "replicated state diverged across ranks")
    torch.cuda.synchronize()
    if rank == 0:
        print(f"DP_GPU_OK world={world} gemm={args.gemm} graph={args.graph} exchange={dp.exchange} steps={steps} worst loss rel={worst['loss']:.2e} w={worst['w']:.2e}", flush=True)
    dp.close(); one.close()
    DADStepEngine.shutdown_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["cpu", "gpu"], required=True)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--T", type=int, default=40)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--gemm", default="fp32")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"])
    a = ap.parse_args()
    run_cpu(a) if a.mode == "cpu" else run_gpu(a)
