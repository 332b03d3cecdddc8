"""Data-parallel path: world_size-2 gloo test of the sharding protocol on CPU, and (gpu) the CUDA engine
at world 2 under NCCL against the same engine at world 1 on the global batch.  Workers: tests/dp_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, extra, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dp_worker.py")] + extra
    env = dict(os.environ, OMP_NUM_THREADS="2", PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


def test_pack_layout_roundtrip():
    """Host logic: the exchange-buffer layout helpers agree with the C ABI's word count and address disjoint fields."""
    from dad_b200 import _lib
    for B, world in ((1, 1), (5, 2), (16, 8)):
        words = _lib.pack_rank_words(B)
        assert words == _lib.lib().dad_pack_rank_words(B) == _lib.pack_offsets(B)["words"]
        pack = torch.zeros(world * words)
        gv = _lib.pack_global_views(pack, B, world)
        for i, k in enumerate(("emb_clean", "emb_strong", "score")):
            gv[k].fill_(float(i + 1))
        gv["pred"].fill_(7); gv["label"].fill_(9)
        seg = pack.view(world, words)
        pi = pack.view(torch.int32).view(world, words)
        assert torch.all(seg[:, :B * 256] == 1) and torch.all(seg[:, B * 256:2 * B * 256] == 2)
        assert torch.all(seg[:, 2 * B * 256:2 * B * 256 + B] == 3)
        assert torch.all(pi[:, 2 * B * 256 + B:2 * B * 256 + 2 * B] == 7) and torch.all(pi[:, 2 * B * 256 + 2 * B:] == 9)
        assert gv["emb_clean"].reshape(world * B, 256).shape == (world * B, 256)


def test_dp_protocol_world2_gloo():
    """Batch-sharded step (1 packed all-gather + 1 all-reduce) == single-process global-batch oracle step; state replicated."""
    r = _torchrun(2, ["--mode", "cpu", "--B", "8", "--T", "24", "--steps", "4"])
    assert r.returncode == 0 and "DP_CPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("gemm,graph,exchange", [("fp32", False, "nccl"), ("tc", True, "nccl"), ("fp32", False, "p2p"), ("tc", True, "p2p")])
def test_dp_engine_world2(gemm, graph, exchange):
    """world 2 == world 1 on the global batch, with the NCCL collectives and with the NVLink peer-memory exchange (k_peer.cu)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = _torchrun(2, ["--mode", "gpu", "--B", "8", "--T", "40", "--steps", "5", "--gemm", gemm, "--exchange", exchange] + (["--graph"] if graph else []))
    assert r.returncode == 0 and "DP_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
