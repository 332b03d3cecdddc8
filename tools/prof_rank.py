"""Per-kernel timeline of ONE data-parallel rank on a single GPU (profiling harness, not a test).

Builds the engine as rank 0 of `world` ranks with the collectives stubbed out and the other ranks' segments of the exchange
buffer pre-filled with plausible embeddings / scores / pseudo-labels / labels, so that every kernel sees exactly the shapes a
real rank sees (global-batch DACP + ECDA statistics, 1/world of the ECDA rows, local encoders).  Run it under
`ncu --metrics gpu__time_duration.sum` for the launch list, or plain for the CUDA-event step time without NCCL.
Usage: python tools/prof_rank.py [world] [global_batch] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import dad_b200  # noqa: F401
from dad_b200 import _lib, config as cfg, synth
from dad_b200.engine import DADStepEngine

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Bg = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 200
B = Bg // world
dist.all_gather_into_tensor = lambda *a, **k: None   # stub: the other segments are pre-filled below
dist.all_reduce = lambda *a, **k: None
dev = torch.device("cuda", 0)
eng = DADStepEngine(cfg, B, T, device=dev, world=world, rank=0, process_group=None, gemm_mode=_lib.GEMM_TC, use_graph=False, noise_mode="philox")
eng.load_state_dict(synth.init_state(42))
g = torch.Generator(device=dev).manual_seed(0)
mu = synth.class_means().to(dev)
y = torch.randint(0, 4, (B,), generator=g, device=dev)
x = mu[y][:, None, :] + torch.randn(B, T, 768, generator=g, device=dev)
xn = x + 0.3 * torch.randn(B, T, 768, generator=g, device=dev)
pad = torch.zeros(B, T, dtype=torch.bool, device=dev)
eng.set_inputs({"net_input": {"feats": x, "padding_mask": pad}, "labels": y}, {"net_input": {"feats": xn, "padding_mask": pad}, "labels": y},
               n_clean_global=Bg)
eng.set_lr(cfg.LEARNING_RATE)
for _ in range(30):
    eng.step(0)
eng.init_teacher()
eng.step(60)
torch.cuda.synchronize()
# other ranks' segments := copies of rank 0's (embeddings jittered), so classes / masks / scores are realistic
gv = _lib.pack_global_views(eng.pack, B, world)
for r in range(1, world):
    gv["emb_clean"][r].copy_(gv["emb_clean"][0] + 0.01 * torch.randn(B, 256, generator=g, device=dev))
    gv["emb_strong"][r].copy_(gv["emb_strong"][0] + 0.01 * torch.randn(B, 256, generator=g, device=dev))
    gv["score"][r].copy_(gv["score"][0]); gv["pred"][r].copy_(gv["pred"][0]); gv["label"][r].copy_(gv["label"][0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    eng.step(60)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    eng.step(60)
e1.record()
torch.cuda.synchronize()
print(f"rank-0-of-{world} step without collectives (eager launches): {e0.elapsed_time(e1) / 10:.3f} ms; nmask={int(eng.nmask)} losses={eng.losses()}")
if os.environ.get("PROF_RANK_GRAPH"):
    # the same step as a replayed CUDA graph (what the bench runs), still without collectives: kernel chain + launch gaps only
    eng.use_graph, eng.graph_after = True, 0
    for _ in range(4):
        eng.step(60)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        eng.step(60)
    e1.record()
    torch.cuda.synchronize()
    print(f"rank-0-of-{world} step without collectives (graph replay): {e0.elapsed_time(e1) / 20:.3f} ms; launches/step {eng.launches_per_step}")
