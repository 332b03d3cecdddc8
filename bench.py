#!/usr/bin/env python3
"""DAD train samples/s on N B200s of one node (+ fraction of roofline, + the reference's CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU implementation (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU (driver launches this for N > 1)

One "step" = one full DAD training step (I/train.py:484-492) on one (clean, noisy) batch pair of synthetic features.
Default workload = BASELINE.json configs[3]: IEMOCAP DAD-train, GLOBAL batch 4096, T=200, post-warm-up epoch 60 (full
loss stack CE + DACP-masked KD + ECDA), sharded over the N ranks (strong scaling; at N=1 the whole 4096 fits one GPU).
`--workload b64` selects configs[1] (batch 64).  One sample = one clean + one noisy utterance."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_IN = 768
FLOPS_PER_PAIR_T = 10 * 768 * 256      # x T: 3 forward + 2 dW GEMMs
BYTES_PER_PAIR_T = 16 * 768            # x T: clean + noisy fp32 features read once forward, once backward


def dominant_kernel(kern, traffic_table=None):
    """The kernel SYMBOL with the largest share of the step and its AVERAGE launch over the launches it has in one step.
    kern: {variant: {"symbol", "ms", "alg_bytes", "alg_flops", "products"}} -- every variant (= one launch of the step: k_enc_fwd_tc
    runs on the clean view and, fused, on the weak + strong views of the noisy batch; k_enc_bwd_tc on the clean and the strong
    view) timed alone.  achieved = algorithmic bytes (flops) of the symbol's launches / their total time; `tensor_TFLOPs` counts
    the EXECUTED products (3 fp16 products per algorithmic FLOP in the forward, 2 in the default backward).  traffic = DRAM bytes
    per launch from the committed ncu capture, averaged the same way (None unless the capture has every variant)."""
    syms = {}
    for name, k in kern.items():
        syms.setdefault(k["symbol"], []).append(name)
    share = {sym: sum(kern[v]["ms"] for v in vs) for sym, vs in syms.items()}
    dom = max(share, key=share.get)
    vs = syms[dom]
    tot = share[dom]
    traffic = None
    if traffic_table and all(traffic_table.get(v) is not None for v in vs):
        traffic = sum(traffic_table[v] for v in vs) / len(vs)
    return {"kernel": dom, "variants": vs, "launches_per_step": len(vs), "avg_launch_ms": tot / len(vs),
            "alg_bytes_per_launch": sum(kern[v]["alg_bytes"] for v in vs) / len(vs),
            "GBps": sum(kern[v]["alg_bytes"] for v in vs) / tot / 1e6, "TFLOPs": sum(kern[v]["alg_flops"] for v in vs) / tot / 1e9,
            "tensor_TFLOPs": sum(kern[v]["alg_flops"] * kern[v]["products"] for v in vs) / tot / 1e9,
            "traffic": traffic, "share_ms": share}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="g4096", choices=["g4096", "b64", "t1000"])
    ap.add_argument("--gemm", default=os.environ.get("DAD_GEMM", "tc"), choices=["tc", "fp32"])
    ap.add_argument("--epoch", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-dp-check", action="store_true", help="N > 1: skip the out-of-timed-region data-parallel correctness check")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the few-step timings of the other BASELINE configs (b64, t1000)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="data-parallel exchange: NVLink peer-memory pushes or NCCL")
    ap.add_argument("--operand-cache", default="strong", choices=["none", "strong", "both"],
                    help="which student forwards write the fp16 hi/lo operand cache that feeds the dW kernels by TMA")
    ap.add_argument("--pretrain-steps", type=int, default=30, help="CE-only steps before the timed region (SURVEY 8d); lower it under ncu")
    return ap.parse_args()


def workload(name, world):
    if name == "g4096":
        return dict(label="IEMOCAP DAD-train global batch 4096 (BASELINE configs[3]), T=200, epoch 60 full loss stack",
                    Bg=4096, T=200, scaling="strong")
    if name == "b64":
        return dict(label="IEMOCAP DAD-train fold 0 batch 64 per GPU (BASELINE configs[1]), T=200, epoch 60 full loss stack",
                    Bg=64 * world, T=200, scaling="weak")
    return dict(label="long-utterance stress B=512 per GPU, T=1000 (BASELINE configs[4])", Bg=512 * world, T=1000, scaling="weak")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (B200_PROFILING.md "clocks line").
    NVML (nvidia-ml-py) gives a sample every ~5 ms; nvidia-smi is the fallback (one sample per ~100 ms)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.sm_max, self.reasons, self.source, self.mask = [], None, set(), "nvidia-smi", 0
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        self.mask |= int(r)
        for name, bit in (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            c = [t.strip() for t in out.split(",")]
            if c[0].replace(".", "").isdigit():
                self.sm.append(float(c[0]))
            self.sm_max = float(c[1])
            self.reasons.update(n for i, n in enumerate(self.NAMES) if c[2 + i] == "Active")

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.004 if self.nvml is not None else 0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"], "samples": 0, "source": self.source}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.sm_max,
                "reasons": [n for n in self.NAMES if n in self.reasons], "samples": len(self.sm), "source": self.source,
                "nvml_reasons_mask": hex(self.mask)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def device_batches(B, T, seed, dev, snr_db=10.0):
    """Synthetic class-structured features generated ON the device (SURVEY.md 8d formulas; the CPU generator in
    dad_b200.synth is used for parity tests, this one for throughput: 2 x B x T x 768 randn would take minutes on the host)."""
    import torch
    from dad_b200 import synth
    g = torch.Generator(device=dev).manual_seed(seed)
    mu = synth.class_means().to(dev)
    out = []
    for domain in ("clean", "noisy"):
        y = torch.randint(0, 4, (B,), generator=g, device=dev)
        x = mu[y][:, None, :] + torch.randn(B, T, D_IN, generator=g, device=dev)
        if domain == "noisy":
            P = (x * x).mean(dim=(1, 2))
            x = x + torch.sqrt(P / (10.0 ** (snr_db / 10.0)))[:, None, None] * torch.randn(B, T, D_IN, generator=g, device=dev)
        out.append({"net_input": {"feats": x.contiguous(), "padding_mask": torch.zeros(B, T, dtype=torch.bool, device=dev)}, "labels": y})
    return out[0], out[1]


def pin_to_gpu_numa_node(dev):
    """Bind this process (and so the pages its pinned staging buffers touch first) to the CPUs of the NUMA node the GPU hangs off:
    with 8 ranks on one box every rank otherwise stages through node 0 (r01: 23 GB/s per rank instead of 54).  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(dev)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def dp_check(cfg, world, rank, dev, pg, B, gm, exchange, T=64, epoch=45):
    """Outside the timed region (N > 1): is the data-parallel step the single-process step?  Every rank builds the SAME global batch
    and RNG draws (same seeds), runs its shard through a world-N engine (supplied-noise mode) and the whole batch through a
    world-1 engine: 30 CE-only steps, teacher := student, 2 full-stack steps.  Reports the bit-identity of the replicated state
    across ranks (all-gathered 64-bit checksum of p, t, ema_thr), the largest relative loss difference and weight difference
    against the world-1 engine, and whether the DACP masks agree (SURVEY.md 8e "Determinism")."""
    import torch
    import torch.distributed as dist
    from dad_b200 import synth
    from dad_b200.engine import DADStepEngine
    Bg = B * world
    clean, noisy = device_batches(Bg, T, 777, dev)
    g = torch.Generator(device=dev).manual_seed(778)
    rng = {"drop_clean": (torch.rand(Bg, 256, generator=g, device=dev) >= 0.1).float(),
           "n_weak": torch.randn(Bg, T, D_IN, generator=g, device=dev), "n_strong": torch.randn(Bg, T, D_IN, generator=g, device=dev),
           "feat_keep_u": torch.rand(D_IN, generator=g, device=dev),
           "t_start": torch.randint(0, T - int(T * 0.1) + 1, (Bg,), generator=g, device=dev),
           "drop_strong": (torch.rand(Bg, 256, generator=g, device=dev) >= 0.1).float()}
    sl = slice(rank * B, (rank + 1) * B)
    sh = lambda b: {"net_input": {k: v[sl].contiguous() for k, v in b["net_input"].items()}, "labels": b["labels"][sl].contiguous()}
    rng_l = {k: (v if k == "feat_keep_u" else v[sl].contiguous()) for k, v in rng.items()}
    dp = DADStepEngine(cfg, B, T, device=dev, world=world, rank=rank, process_group=pg, gemm_mode=gm, use_graph=True,
                       noise_mode="supplied", exchange=exchange)
    one = DADStepEngine(cfg, Bg, T, device=dev, gemm_mode=gm, use_graph=False, noise_mode="supplied")
    st = synth.init_state(42)
    loss_rel = 0.0
    for e in (dp, one):
        e.load_state_dict(st)
        e.set_lr(cfg.LEARNING_RATE)
    n_ce = 30  # CE-only steps first: the teacher must be confident for DACP to mask samples in and KD / ECDA to run
    for k in range(n_ce + 2):
        ep = 0 if k < n_ce else epoch
        if k == n_ce:
            dp.init_teacher(); one.init_teacher()
        one.set_inputs(clean, noisy, rng)
        dp.set_inputs(sh(clean), sh(noisy), rng_l)
        one.step(ep); dp.step(ep)
        a, b = dp.losses(), one.losses()
        for key in a:
            loss_rel = max(loss_rel, abs(a[key] - b[key]) / max(abs(b[key]), 1e-3))
    mask_equal = bool(torch.equal(dp.mask, one.mask))
    w_diff = float(torch.maximum((dp.p - one.p).abs().max(), (dp.t - one.t).abs().max()))
    well = one.g_ext[:dp.p.numel()].abs() >= 1e-6  # elements whose Adam update is well conditioned (as in the parity tests)
    w_diff_well = float((dp.p - one.p).abs()[well].max()) if bool(well.any()) else 0.0
    flat = torch.cat([dp.p, dp.t, dp.ema_thr]).view(torch.int32).to(torch.int64)
    h = (flat * (torch.arange(flat.numel(), device=dev) % 65521 + 1)).sum().reshape(1)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h, group=pg)
    out = {"state_identical": bool(all(int(x) == int(hs[0]) for x in hs)), "loss_rel": loss_rel, "weights_max_abs_diff_vs_world1": w_diff,
           "weights_max_abs_diff_well_conditioned": w_diff_well, "last_losses": a,
           "mask_equal": mask_equal, "masked_in": int(dp.nmask.item()), "B_local": B, "B_global": Bg, "T": T, "steps": f"{n_ce} CE-only + 2 full",
           "exchange": dp.exchange}
    dp.close(); one.close()
    return out


def time_small_workload(cfg, dev, gm, B, T, epoch, steps):
    """A few device-timed steps (graph replay, CUDA events) of another BASELINE config on this GPU: ms per step."""
    import torch
    from dad_b200 import synth
    from dad_b200.engine import DADStepEngine
    eng = DADStepEngine(cfg, B, T, device=dev, gemm_mode=gm, use_graph=True, noise_mode="philox", seed=42)
    eng.load_state_dict(synth.init_state(42))
    clean, noisy = device_batches(B, T, 4321, dev)
    eng.set_inputs(clean, noisy)
    eng.set_lr(cfg.LEARNING_RATE)
    for _ in range(10):
        eng.step(0)
    eng.init_teacher()
    for _ in range(4):
        eng.step(epoch)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0.record(); eng.step(epoch); e1.record()
        torch.cuda.synchronize(dev)
        tot += e0.elapsed_time(e1)
    out = {"ms_per_step": tot / steps, "samples_per_s": B / (tot / steps * 1e-3), "B": B, "T": T, "steps": steps,
           "launches_per_step": eng.launches_per_step, "l2": "flushed between individually timed steps"}
    eng.close()
    del eng, clean, noisy
    torch.cuda.empty_cache()
    return out


def time_predict(dev, B, T, reps=5):
    """Validation pass (I/train.py:522-564) over one [B,T,768] batch: student + teacher predictions from ONE fused read of x
    (dad_predict) against two single-network passes; CUDA events."""
    import torch
    from dad_b200 import ops, synth
    st = synth.init_state(42)
    cu = lambda k: st[k].to(dev)
    Ws, bs, Wc, bc = cu("student_encoder.pre_net.weight"), cu("student_encoder.pre_net.bias"), cu("student_classifier.fc_layer.weight"), cu("student_classifier.fc_layer.bias")
    x = torch.randn(B, T, D_IN, device=dev)
    split = ops.split_f16(Ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps
    both = timed(lambda: ops.predict(x, None, Ws, bs, Wc, bc, Ws, bs, Wc, bc, Ws_split=split, Wt_split=split))
    one = timed(lambda: ops.predict(x, None, Ws, bs, Wc, bc, Ws_split=split))
    return {"B": B, "T": T, "student_and_teacher_one_pass_ms": both, "one_network_ms": one, "two_separate_passes_ms": 2 * one}


def cpu_reference_steps(B, T, steps, warmup, epoch, threads=None):
    """The reference's CPU step, restated op for op in oracle/dad_oracle.py (the Python reference cannot travel to the
    GPU box; the oracle is bit-identical to it on CPU, tests/test_oracle_golden.py).  Returns seconds per step (median)."""
    import torch
    from dad_b200 import config as cfg, synth
    from oracle import dad_oracle as O
    torch.set_num_threads(threads or os.cpu_count())
    hp = O.HParams.from_cfg(cfg)
    st = synth.init_state(42)
    orc = O.OracleStepper(st, hp)
    c, n = synth.make_batches(B, T, 1)
    gen = torch.Generator().manual_seed(2)
    for i in range(30):  # CE-only "pretraining" so the teacher is confident and every branch executes
        orc.step(c, n, 0, synth.make_rng_inputs(B, T, 1000 + i))
    for ks, kt in zip(O.PARAM_KEYS, O.TEACHER_KEYS):
        orc.t[kt] = orc.p[ks].detach().clone()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        # the reference draws its noise inside the step (torch.randn_like etc., I/utils.py:330-371): keep that cost inside
        rng = {"drop_clean": (torch.rand(B, 256, generator=gen) >= 0.1).float(), "n_weak": torch.randn(B, T, D_IN, generator=gen),
               "n_strong": torch.randn(B, T, D_IN, generator=gen), "feat_keep_u": torch.rand(D_IN, generator=gen),
               "t_start": torch.randint(0, T - int(T * 0.1) + 1, (B,), generator=gen),
               "drop_strong": (torch.rand(B, 256, generator=gen) >= 0.1).float()}
        orc.step(c, n, epoch, rng)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return statistics.median(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload, args.gpus)
    # each step = ONE bounded sample of the workload: a B=64 sub-batch (the reference's own batch size, I/config.py:36; its ECDA
    # materialises [n,n,256] differences, I/utils.py:537 -- 4096 at once would need hundreds of GB).  --steps / --warmup are honoured
    # as given: a B=64, T=200 step takes ~0.13 s on 16 cores, so the driver's 20 + 5 finish in seconds.
    Bs, T = 64, wl["T"]
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    sec, threads = cpu_reference_steps(Bs, T, steps, warmup, args.epoch)
    val = Bs / sec
    sample = f"each step = one B={Bs}, T={T} sub-batch of the workload (full loss stack, epoch {args.epoch}); {steps} timed steps, median"
    print(json.dumps({
        "impl": "reference", "metric": "DAD train samples/sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "global_batch": wl["Bg"], "T": T, "device": "host CPU", "sub_batch": Bs},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import dad_b200  # noqa: F401
    from dad_b200 import _lib, config as cfg, synth
    from dad_b200.engine import DADStepEngine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = pin_to_gpu_numa_node(dev)
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    wl = workload(args.workload, world)
    Bg, T = wl["Bg"], wl["T"]
    assert Bg % world == 0
    B = Bg // world
    gm = _lib.GEMM_TC if args.gemm == "tc" else _lib.GEMM_FP32
    eng = DADStepEngine(cfg, B, T, device=dev, world=world, rank=rank, process_group=pg, gemm_mode=gm, use_graph=not args.no_graph,
                        noise_mode="philox", seed=42, operand_cache=args.operand_cache, exchange=args.exchange)
    eng.load_state_dict(synth.init_state(42))
    clean, noisy = device_batches(B, T, 1234 + rank, dev)
    eng.set_inputs(clean, noisy)
    eng.set_lr(cfg.LEARNING_RATE)
    for _ in range(args.pretrain_steps):  # CE-only pretraining (epoch 0 < WARMUP_EPOCHS), then teacher := student, fresh Adam
        eng.step(0)
    eng.init_teacher()
    eng.m.zero_(); eng.v.zero_(); eng.step_count.zero_()
    eng.set_lr(eng.cosine_lr(args.epoch))
    sync = lambda: (torch.cuda.synchronize(dev), world > 1 and torch.distributed.barrier())

    # ---- device-timed steps: inputs resident in HBM
    for _ in range(max(args.warmup, 3)):
        eng.step(args.epoch)
    sync()
    if eng.exchange == "p2p":
        eng._sym[16:64].zero_()  # wait statistics of the flag barriers (csrc/dad_optim.cuh peer_flag_barrier)
        torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    sampler.start()
    small = 2 * B * T * D_IN * 4 < 2 * 126e6  # working set not >> L2: flush L2 between steps and time steps individually
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if small else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if small:
        tot = 0.0
        for _ in range(args.steps):
            flush.zero_()
            e0.record(); eng.step(args.epoch); e1.record()
            torch.cuda.synchronize(dev)
            tot += e0.elapsed_time(e1)
        ms = tot / args.steps
    else:
        e0.record()
        th0 = time.perf_counter()
        for _ in range(args.steps):
            eng.step(args.epoch)
        host_enqueue_ms = (time.perf_counter() - th0) * 1e3 / args.steps  # host time to enqueue one step (no sync inside the loop)
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.summary()
    ms_own = ms
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    barrier_wait = None
    host_enqueue = None
    if world > 1 and not small:
        he = torch.tensor([host_enqueue_ms, ms_own], device=dev, dtype=torch.float64)
        allh = [torch.zeros_like(he) for _ in range(world)]
        torch.distributed.all_gather(allh, he)
        host_enqueue = {"host_ms_per_step": [round(float(a[0]), 4) for a in allh], "device_ms_per_step": [round(float(a[1]), 4) for a in allh]}
    if eng.exchange == "p2p":
        # average microseconds per step that rank i (row) waited for rank j (column) at the pack / gradient barrier: the row maximum
        # is what the barrier cost that rank, a column of large values names a late rank
        w = eng._sym[32:64].view(torch.int64).to(torch.float64).reshape(2, 8)[:, :world] / (1e3 * args.steps)
        allw = [torch.zeros_like(w) for _ in range(world)]
        torch.distributed.all_gather(allw, w.contiguous())
        pd = eng._sym[16:32].view(torch.int64)[1:2].to(torch.float64) / (1e3 * args.steps)  # push duration (k_peer_push_barrier)
        allp = [torch.zeros_like(pd) for _ in range(world)]
        torch.distributed.all_gather(allp, pd.contiguous())
        barrier_wait = {"gradient_us": [[round(float(v), 1) for v in a[0]] for a in allw],
                        "pack_us": [[round(float(v), 1) for v in a[1]] for a in allw],
                        "pack_push_us": [round(float(a[0]), 1) for a in allp]}
    launches = eng.launches_per_step * args.steps
    losses = eng.losses()
    value = Bg / (ms * 1e-3)
    cfg_snapshot = {"operand_cache": eng.operand_cache, "exchange": eng.exchange, "fused_noisy": eng.fused_noisy,
                    "cta_pair": bool(_lib.lib().dad_tc_get_cta_pair()), "bwd_products": eng.bwd_products}
    dtype_s = "f32" if args.gemm != "tc" else (f"f32 (scaled fp16 hi/lo split: 3 tcgen05 products per forward GEMM, {eng.bwd_products} per dW GEMM, fp32 accumulate)")

    # ---- the encoder launches of one step, each timed ALONE (roofline).  Every variant is captured into its own CUDA graph and the
    # graph is replayed under CUDA events, so the figure is device time of the launcher's kernels (GEMM kernel + its helpers:
    # k_pool_finalize, or absmax / prep / db / fold) without host enqueue gaps.
    peaks, peak_src = measured_peaks()
    sc = eng.scalars(args.epoch)
    reps = 10 if not small else 40
    aug_w, aug_s = eng._aug(1, sc), eng._aug(2, sc)
    emb_c, emb_s = eng.pack.data_ptr(), eng.pack.data_ptr() + B * 256 * 4
    view_bytes, gemm_flops = B * T * D_IN * 4, 2 * B * T * D_IN * 256
    P = eng.bwd_products if args.gemm == "tc" else 1
    variants = [("encoder_fwd_clean", "k_enc_fwd_tc", view_bytes, gemm_flops, 3,
                 lambda: eng._enc_fwd(eng.x_clean, eng.pad_clean, False, None, emb_c, eng.bits_clean, eng.lens_clean, _lib.stream_ptr(), xs=eng.xs_clean))]
    if eng.fused_noisy:
        fused_symbol = "k_enc_fwd_tp" if _lib.lib().dad_tc_get_cta_pair() == 2 else "k_enc_fwd_tc"  # default policy: the transposed CTA-pair kernel
        variants.append(("encoder_fwd_noisy_fused", fused_symbol, view_bytes, 2 * gemm_flops, 3,
                         lambda: eng._enc_fwd_noisy_fused(aug_w, aug_s, emb_s, _lib.stream_ptr())))
    else:
        variants += [("encoder_fwd_weak", "k_enc_fwd_tc", view_bytes, gemm_flops, 3,
                      lambda: eng._enc_fwd(eng.x_noisy, eng.pad_noisy, True, aug_w, eng.emb_teacher.data_ptr(), None, eng.lens_scratch, _lib.stream_ptr())),
                     ("encoder_fwd_strong", "k_enc_fwd_tc", view_bytes, gemm_flops, 3,
                      lambda: eng._enc_fwd(eng.x_noisy, eng.pad_noisy, False, aug_s, emb_s, eng.bits_strong, eng.lens_noisy, _lib.stream_ptr(), xs=eng.xs_strong))]
    variants += [("encoder_bwd_clean", "k_enc_bwd_tc", view_bytes, gemm_flops, P,
                  lambda: eng._enc_bwd(eng.x_clean, eng.pad_clean, None, eng.bits_clean, eng.g_clean, eng.lens_clean, 0, _lib.stream_ptr(), xs=eng.xs_clean)),
                 ("encoder_bwd_strong", "k_enc_bwd_tc", view_bytes, gemm_flops, P,
                  lambda: eng._enc_bwd(eng.x_noisy, eng.pad_noisy, aug_s, eng.bits_strong, eng.g_strong, eng.lens_noisy, 1, _lib.stream_ptr(), xs=eng.xs_strong))]
    kern = {}
    for name, symbol, nbytes, nflops, products, fn in variants:
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize(dev)
        tot = 0.0
        for _ in range(reps):
            if flush is not None:
                flush.zero_()
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize(dev)
            tot += e0.elapsed_time(e1)
        t_ms = tot / reps
        kern[name] = {"symbol": symbol if args.gemm == "tc" else symbol.replace("_tc", "_simt"), "ms": t_ms, "alg_bytes": nbytes, "alg_flops": nflops,
                      "products": products if args.gemm == "tc" else 1, "GBps": nbytes / t_ms / 1e6, "TFLOPs": nflops / t_ms / 1e9}
        del g
    # dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` capture of this workload at this per-GPU
    # batch (profiles/traffic.json; null when no capture matches)
    table = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            table = json.load(f).get(f"{args.workload}:B{B}:T{T}:{args.operand_cache}:r02")
    except Exception:
        pass
    dk = dominant_kernel(kern, table)
    # executed tensor work of one step: 3 forward GEMMs x 3 products + 2 dW GEMMs x P products
    exec_flops_step = (3 * 3 + 2 * P) * gemm_flops if args.gemm == "tc" else 5 * gemm_flops
    roofline = {"bound": "hbm", "kernel": dk["kernel"], "variants": dk["variants"], "launches_per_step": dk["launches_per_step"],
                "avg_launch_ms": dk["avg_launch_ms"], "achieved": dk["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": dk["GBps"] / peaks["hbm_gbs"], "traffic": dk["traffic"], "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dk["alg_bytes_per_launch"], "kernels": kern,
                "note": ("the dominant launch is the fused noisy forward: ONE read of x_noisy (its algorithmic bytes) feeds two GEMMs (weak -> teacher, "
                         "strong -> student), so its HBM fraction is low by construction and `tensor_3pass` is the roofline that binds it"
                         if dk["kernel"] == "k_enc_fwd_tp" else None),
                # the same launches seen from the tensor pipe (executed fp16 products; fp32-grade accuracy costs 3 per algorithmic FLOP in
                # the forward), against the BURST peak: these kernels are timed in isolation
                "tensor_3pass": {"achieved": dk["tensor_TFLOPs"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                 "frac": dk["tensor_TFLOPs"] / peaks["bf16_tflops"], "peak_kind": "burst (kernel timed alone)"} if args.gemm == "tc" else None,
                # the whole step, inside a long timed loop: against the SUSTAINED tensor peak
                "step": {"algorithmic_bytes": BYTES_PER_PAIR_T * T * B, "algorithmic_flops": FLOPS_PER_PAIR_T * T * B,
                         "hbm_roofline_ms": BYTES_PER_PAIR_T * T * B / (peaks["hbm_gbs"] * 1e6),
                         "frac_of_hbm_roofline": (BYTES_PER_PAIR_T * T * B / (peaks["hbm_gbs"] * 1e6)) / ms,
                         "executed_tensor_flops": exec_flops_step, "bwd_products": P,
                         "tensor_roofline_ms": exec_flops_step / (peaks["bf16_tflops_sustained"] * 1e9),
                         "frac_of_tensor_roofline": exec_flops_step / (peaks["bf16_tflops_sustained"] * 1e9) / ms,
                         "tensor_peak_kind": "sustained (inside the timed step loop)"}}

    # ---- end to end through the public API with HOST buffers (pinned), H2D of the step's inputs + D2H of its losses inside
    e2e = None
    e2e_resident = None
    if not args.no_e2e:
        hb = []
        for b in (clean, noisy):
            hb.append({"net_input": {"feats": b["net_input"]["feats"].cpu().pin_memory(),
                                     "padding_mask": b["net_input"]["padding_mask"].cpu().pin_memory()},
                       "labels": b["labels"].cpu().pin_memory()})
        h2d = sum(t.numel() * t.element_size() for b in hb for t in (b["net_input"]["feats"], b["net_input"]["padding_mask"], b["labels"]))
        n_e2e = max(3, min(args.steps, 10))
        # pipelined public API: the H2D copy of step k+1's batches (copy stream, into staging buffers) overlaps step k; every
        # step's inputs still cross PCIe inside the timed region and every step's losses are read back
        for _ in range(2):
            eng.prefetch_inputs(hb[0], hb[1]); eng.commit_prefetched(); eng.step(args.epoch); eng.losses()
        sync()
        t0 = time.perf_counter()
        eng.prefetch_inputs(hb[0], hb[1])
        for i in range(n_e2e):
            eng.commit_prefetched()
            eng.step(args.epoch)
            if i + 1 < n_e2e:
                eng.prefetch_inputs(hb[0], hb[1])
            eng.losses()  # D2H read of the step's losses (sync)
        sync()
        dt_local = (time.perf_counter() - t0) / n_e2e
        dt = dt_local
        h2d_rates = [h2d / dt_local / 1e9]
        if world > 1:
            t = torch.tensor([dt], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dt = float(t)
            rates = [torch.zeros(1, device=dev) for _ in range(world)]
            torch.distributed.all_gather(rates, torch.tensor([h2d / dt_local / 1e9], device=dev))
            h2d_rates = [float(r) for r in rates]
        e2e = {"value": Bg / dt, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 20, "ms_per_step": dt * 1e3,
               "h2d_GBps_per_rank": h2d_rates, "numa_node_of_rank0": numa_node}
        del hb
        # the trainers' default path (cfg.DAD_DEVICE_DATA): the split's utterances stay in HBM and every batch is assembled ON the
        # device straight into the engine's input buffers (dataload.DeviceUtteranceLoader -> dad_gather_pad) -- per step one gather
        # of each domain's batch, the step, and the D2H read of its losses; no host->device feature traffic at all
        from dad_b200 import dataload
        ld_c = dataload.DeviceUtteranceLoader.from_device_batch(clean, batch_size=B, sink=lambda b_, t_: eng.input_views("clean", b_, t_))
        ld_n = dataload.DeviceUtteranceLoader.from_device_batch(noisy, batch_size=B, sink=lambda b_, t_: eng.input_views("noisy", b_, t_), use_labels=False)
        ids = torch.arange(B)
        for _ in range(2):
            eng.set_inputs(ld_c.collate(ids), ld_n.collate(ids)); eng.step(args.epoch); eng.losses()
        sync()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            eng.set_inputs(ld_c.collate(ids), ld_n.collate(ids))
            eng.step(args.epoch)
            eng.losses()
        sync()
        dtr = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dtr], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dtr = float(t)
        e2e_resident = {"value": Bg / dtr, "unit": "samples/s", "ms_per_step": dtr * 1e3, "h2d_bytes_per_step": 2 * B * 8,
                        "d2h_bytes_per_step": 20, "path": "DeviceUtteranceLoader.collate (dad_gather_pad into the engine's buffers) -> set_inputs -> step -> losses()"}
        del ld_c, ld_n

    # ---- N > 1: correctness signal of the data-parallel step, outside the timed region
    dpc = None
    if world > 1 and not args.no_dp_check:
        eng.close()
        dpc = dp_check(cfg, world, rank, dev, pg, B, gm, args.exchange)
    # ---- the other single-GPU BASELINE configs, a few steps each (driver-run numbers for configs[1] and configs[4])
    extra = None
    if world == 1 and rank == 0 and not args.no_extra and args.workload == "g4096":
        eng.close()
        del eng
        torch.cuda.empty_cache()
        extra = {"b64": time_small_workload(cfg, dev, gm, 64, 200, args.epoch, 20),
                 "t1000": time_small_workload(cfg, dev, gm, 512, 1000, args.epoch, 5),
                 "validate": time_predict(dev, 4096, 200)}
        eng = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, threads = cpu_reference_steps(64, T, 8 if T <= 200 else 2, 2, args.epoch)
        cpu = {"value": 64 / sec, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"oracle port of the reference step, B=64, T={T}, full loss stack, median of {8 if T <= 200 else 2} steps"}
    if rank == 0:
        exch = dpc["exchange"] if dpc else ("none" if world == 1 else args.exchange)
        print(json.dumps({
            "metric": "DAD train samples/sec", "value": value, "unit": "samples/s", "utterances_per_s": 2 * value, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": dtype_s, "data": "synthetic",
            "config": {"workload": wl["label"], "global_batch": Bg, "per_gpu_batch": B, "T": T, "parallelism": f"dp{world}",
                       "noise": "in-kernel Philox", "cuda_graph": not args.no_graph, "operand_cache": cfg_snapshot["operand_cache"],
                       "exchange": cfg_snapshot["exchange"], "fused_noisy_forward": cfg_snapshot["fused_noisy"],
                       "cta_pair": cfg_snapshot["cta_pair"], "bwd_products": cfg_snapshot["bwd_products"],
                       "l2": "inputs >> L2 (126 MB)" if not small else "L2 flushed (256 MB write) between individually timed steps"},
            "clocks": clocks, "e2e": e2e, "e2e_resident": e2e_resident, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "losses_last_step": losses, "dp_check": dpc, "barrier_wait": barrier_wait, "per_rank": host_enqueue, "extra_workloads": extra,
        }))
    if world > 1:
        if eng is not None:
            eng.close()
        DADStepEngine.shutdown_process_group()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner from C) write to fd 1, so fd 1 is pointed at
    stderr for the rest of the process and the JSON goes to a private duplicate of the original stdout."""
    global print
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(keep, "w")
    import builtins

    def _print(*args, **kw):
        kw.setdefault("file", out)
        builtins.print(*args, **kw)
        kw["file"].flush()
    print = _print


if __name__ == "__main__":
    a = parse()
    _claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
