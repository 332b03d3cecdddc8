"""Micro-benchmark of the data-parallel exchange kernels alone (run under torchrun, one rank per GPU):
   pack all-gather (dad_peer_push_barrier), gradient all-reduce (dad_grad_reduce_xchg + dad_grad_norm_xchg) against the rank-local
   dad_grad_reduce + dad_grad_norm, and the NCCL collectives of the same payloads.  All ranks loop in lock-step, so the per-iteration
   time is transfer + barrier round trip without inter-rank skew.
Usage: torchrun --nproc-per-node N tools/xchg_bench.py [B_local] [T] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import dad_b200  # noqa: F401
from dad_b200 import _lib, config as cfg, synth
from dad_b200._lib import NPARAM
from dad_b200.engine import DADStepEngine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
IT = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
eng = DADStepEngine(cfg, B, T, device=dev, world=world, rank=rank, process_group=dist.group.WORLD, gemm_mode=_lib.GEMM_TC, use_graph=False,
                    noise_mode="philox", exchange="p2p")
eng.load_state_dict(synth.init_state(42))
L = _lib.lib()
s = _lib.stream_ptr()
S = L.dad_encoder_bwd_splits(B, T, _lib.GEMM_TC)
rwb = eng._rw * 4
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, name):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(IT):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / IT * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:48s} {float(t):8.1f} us")


def push():
    _lib.check(L.dad_peer_push_barrier(eng.pack.data_ptr() + rank * rwb, eng._peer_ptrs, world, rank, eng._pack_off + rank * rwb, rwb, 0,
                                       eng._peer_seq.data_ptr(), eng._xchg_counters.data_ptr(), s))


def barrier_only():
    _lib.check(L.dad_peer_barrier(eng._peer_ptrs, world, rank, 0, eng._peer_seq.data_ptr(), s))


dWc = eng.dWc_dbc.data_ptr()
dbc = dWc + 4 * 256 * 4


def reduce_xchg():
    _lib.check(L.dad_grad_reduce_xchg(eng._dWp[0].data_ptr(), eng._dbp[0].data_ptr(), S, eng._dWp[1].data_ptr(), eng._dbp[1].data_ptr(), S, dWc, dbc,
                                      eng.g_ext.data_ptr(), eng._peer_ptrs, world, rank, eng._stage_off, 0, eng._peer_seq.data_ptr(),
                                      eng._xchg_counters.data_ptr() + 4, s))
    _lib.check(L.dad_grad_norm_xchg(eng.g_ext.data_ptr(), eng._stage.data_ptr(), world, eng.norm_part.data_ptr(), eng.counters.data_ptr() + 16,
                                    eng.step_count.data_ptr(), eng.lr.data_ptr(), eng.scal.data_ptr(), 1.0, s))


def reduce_local():
    _lib.check(L.dad_grad_reduce(eng._dWp[0].data_ptr(), eng._dbp[0].data_ptr(), S, eng._dWp[1].data_ptr(), eng._dbp[1].data_ptr(), S, dWc, dbc,
                                 eng.g_ext.data_ptr(), s))
    _lib.check(L.dad_grad_norm(eng.g_ext.data_ptr(), eng.norm_part.data_ptr(), eng.counters.data_ptr() + 16, eng.step_count.data_ptr(),
                               eng.lr.data_ptr(), eng.scal.data_ptr(), 1.0, s))


if rank == 0:
    print(f"world {world}, B_local {B}: pack segment {rwb / 1e6:.2f} MB per rank, gradient {(NPARAM + 8) * 4 / 1e6:.2f} MB, {S} split-K partials per view")
timed(barrier_only, "flag barrier alone (dad_peer_barrier)")
timed(push, "pack all-gather (dad_peer_push_barrier)")
timed(reduce_local, "rank-local dad_grad_reduce + dad_grad_norm")
timed(reduce_xchg, "dad_grad_reduce_xchg + dad_grad_norm_xchg")
seg = eng.pack[rank * eng._rw:(rank + 1) * eng._rw]
timed(lambda: dist.all_gather_into_tensor(eng.pack, seg), "NCCL all_gather_into_tensor (pack)")
timed(lambda: dist.all_reduce(eng.g_ext), "NCCL all_reduce (gradient)")
eng.close()
DADStepEngine.shutdown_process_group()
