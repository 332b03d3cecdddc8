"""CPU: the C-ABI library loads and exports every symbol include/dad_sm100a.h declares, the host-side schedule / config mirror
matches the reference, and nothing falls back to the CPU."""
import ctypes
import json
import os
import re

import pytest
import torch

import dad_b200  # noqa: F401
from dad_b200 import _lib, config, config_casia, config_emodb
from dad_b200.engine import StepScalars
from helpers import GOLDEN, cfg_for
from oracle import dad_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dad_sm100a.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:void|int|size_t|int32_t|int64_t|const char\*)\s+(dad_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 25 and "dad_encoder_fwd" in names and "dad_peer_push" in names
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= set(names), sorted(set(_lib.EXPORTS) - set(names))
    assert _lib.lib().dad_abi_version() == 3 and _lib.lib().dad_max_world() == 16
    assert _lib.lib().dad_pack_rank_words(7) == 7 * (2 * 256 + 3)


def test_no_cpu_fallback():
    from dad_b200 import ops
    from dad_b200.engine import DADStepEngine
    if torch.cuda.is_available():
        pytest.skip("this check is about hosts without a GPU")
    with pytest.raises(_lib.DadError):
        ops.encoder_pool_fwd(torch.zeros(2, 4, 768), None, torch.zeros(256, 768), torch.zeros(256))
    with pytest.raises(_lib.DadError):
        ops.augment(torch.zeros(2, 4, 768), 0.01, seed=1)
    with pytest.raises(_lib.DadError):
        DADStepEngine(config, 4, 32)
    from dad_b200.train import IEMOCAPCrossDomainTrainer
    with pytest.raises(_lib.DadError):
        IEMOCAPCrossDomainTrainer(fold=0, loaders=lambda bs: None, require_pretrained=False)


@pytest.mark.parametrize("dataset", ["iemocap", "casia", "emodb"])
def test_step_scalars_follow_the_reference_schedule(dataset):
    cfg = cfg_for(dataset)
    hp = O.HParams.from_cfg(cfg)
    for epoch in (0, 5, 29, 30, 31, 45, 60, 90, 499):
        sc = StepScalars(cfg, epoch, dataset)
        w_kd, w_ecda = O.loss_weights(epoch, hp)  # I/train.py:380-395
        assert sc.warmup == (epoch < cfg.WARMUP_EPOCHS) and sc.w_kd == w_kd and sc.w_ecda == w_ecda
        g = cfg.DACP_QUANTILE_START + (cfg.DACP_QUANTILE_END - cfg.DACP_QUANTILE_START) * (epoch / cfg.EPOCHS)
        assert abs(sc.gamma - g) < 1e-7 and sc.w_scl == 0.0
    sc = StepScalars(cfg, 60, dataset)
    if dataset == "emodb":   # E/train_emodb.py:418-437 has no USE_DACP / USE_ECDA switches
        assert sc.use_dacp and sc.use_ecda
    elif dataset == "casia":  # C/config_casia.py defaults
        assert not sc.use_dacp and not sc.use_ecda
    else:
        assert sc.use_dacp and sc.use_ecda and sc.use_entropy and sc.class_aware
    over = cfg_for(dataset, {"USE_ENTROPY_IN_SCORE": False, "USE_CLASS_AWARE_MMD": False, "USE_DACP": True})
    sc = StepScalars(over, 60, dataset)  # only I/utils.py honours these two switches
    assert sc.use_entropy == (dataset != "iemocap") and sc.class_aware == (dataset != "iemocap")


def test_config_modules_keep_the_reference_field_names_and_values():
    golden = json.load(open(os.path.join(GOLDEN, "config_fields.json")))
    for tag, mod in (("iemocap", config), ("casia", config_casia), ("emodb", config_emodb)):
        for name, want in golden[tag].items():
            assert hasattr(mod, name), (tag, name)
            if want != "<non-scalar or path>" and name not in ("DEVICE",):
                got = getattr(mod, name)
                assert got == want, (tag, name, got, want)
        for fn in ("validate_config", "print_config", "setup_environment"):
            assert callable(getattr(mod, fn))


def test_runner_style_override_is_seen_live():
    """The ablation runners setattr on the config module after import (I/run_granular_ablations.py:25-44): the step reads live."""
    old = config.USE_ECDA
    try:
        config.USE_ECDA = False
        assert not StepScalars(config, 60, "iemocap").use_ecda
        config.USE_ECDA = True
        assert StepScalars(config, 60, "iemocap").use_ecda
    finally:
        config.USE_ECDA = old


def test_ctypes_mirrors_match_the_c_structs(tmp_path):
    """include/dad_sm100a.h is plain C: compile it with gcc (no CUDA headers needed) and compare sizeof / offsetof of every
    argument struct with the ctypes mirrors in dad_b200/_lib.py, field by field and by NAME (a field renamed, reordered or
    re-typed on one side only would otherwise surface as silent garbage in a kernel argument)."""
    import ctypes
    import subprocess
    from dad_b200 import _lib
    pairs = [("dad_aug", _lib.Aug), ("dad_pack", _lib.Pack), ("dad_head_fwd_args", _lib.HeadFwdArgs), ("dad_dacp_args", _lib.DacpArgs),
             ("dad_head_bwd_args", _lib.HeadBwdArgs), ("dad_ecda_args", _lib.EcdaArgs), ("dad_scl_args", _lib.SclArgs)]
    inc = os.path.join(ROOT, "include")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dad_sm100a.h"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'    printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'    printf("{cname}.{fname} %zu %zu\\n", offsetof({cname}, {fname}), sizeof((({cname}*)0)->{fname}));')
    lines += ['    printf("version %d nparam %d world %d\\n", DAD_ABI_VERSION, DAD_NPARAM, DAD_MAX_WORLD);', "    return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = {}
    for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines():
        k, *v = ln.split()
        got[k] = v
    for cname, cls in pairs:
        assert int(got[cname][0]) == ctypes.sizeof(cls), (cname, got[cname], ctypes.sizeof(cls))
        for fname, ftype in cls._fields_:
            off, size = (int(t) for t in got[f"{cname}.{fname}"])
            assert off == getattr(cls, fname).offset and size == ctypes.sizeof(ftype), (cname, fname, off, size)
    assert got["version"] == ["3", "nparam", str(_lib.NPARAM), "world", "16"]


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/dad_sm100a.h against the argtypes / restype dad_b200/_lib.py declares: same arity, and the same
    kind (pointer / int32 / int64 / uint64 / size_t / float) in every position."""
    L = _lib.lib()
    src = open(os.path.join(ROOT, "include", "dad_sm100a.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(int|size_t|int32_t|int64_t|const char\*)\s+(dad_\w+)\s*\(([^)]*)\)\s*;", src)
    assert len(protos) >= 25
    scalar = {"int": ctypes.c_int32, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
              "size_t": ctypes.c_size_t, "float": ctypes.c_float}
    checked = 0
    for ret, name, params in protos:
        fn = getattr(L, name)
        plist = [] if params.strip() in ("", "void") else [q.strip() for q in params.split(",")]
        if fn.argtypes is None:
            # undeclared argtypes are only safe for functions without pointer / 64-bit / float parameters
            assert all("*" not in q and q.split()[0] in ("int", "int32_t") for q in plist), (name, params)
            continue
        assert len(fn.argtypes) == len(plist), (name, len(fn.argtypes), plist)
        for i, (q, at) in enumerate(zip(plist, fn.argtypes)):
            if "*" in q:
                ok = at is ctypes.c_void_p or isinstance(at, type) and issubclass(at, ctypes._Pointer)
            else:
                base = q.replace("const", "").split()[0]
                ok = at is scalar[base] or (ctypes.sizeof(at) == ctypes.sizeof(scalar[base]) and at._type_ == scalar[base]._type_)
            assert ok, (name, i, q, at)
        if ret == "size_t":
            assert fn.restype is ctypes.c_size_t, name
        if ret == "int64_t":
            assert fn.restype is ctypes.c_int64, name
        checked += 1
    assert checked >= 20


def test_tensor_core_kernels_are_tcgen05_and_tma_in_sass():
    """Static evidence from the built library (cuobjdump, no GPU): the four GEMM kernels issue tcgen05.mma (UTCHMMA) with TMEM
    read-back (LDTM), commit to mbarriers (UTCBAR) and are fed by TMA (UTMALDG; the encoder forward also TMA-stores its operand
    cache, UTMASTG); no kernel falls back to legacy mma.sync (HMMA); the big kernels keep their registers (no spill stack)."""
    import importlib.util
    import shutil
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")) or not shutil.which("c++filt"):
        pytest.skip("cuobjdump / c++filt not available")
    spec = importlib.util.spec_from_file_location("sass_summary", os.path.join(ROOT, "tools", "sass_summary.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    s = mod.summary()
    for base in ("k_enc_fwd_tc", "k_enc_bwd_tc", "k_gram_tc", "k_pz_tc"):
        inst = [k for k in s if k.split("<")[0] == base]  # every template instance (views x CTA pair, products)
        assert inst, base
        for k in inst:
            assert s[k]["UTCHMMA"] > 0 and s[k]["UTCBAR"] > 0 and s[k]["LDTM"] > 0 and s[k]["UTMALDG"] > 0, (k, dict(s[k]))
            # spill stack: none in the GEMM-only kernels; the forward's converter warps run at 88 registers (setmaxnreg budget of a
            # 896-thread CTA) and keep a few words of per-tile state in L1-resident local memory
            assert s[k]["stack"] <= (96 if base == "k_enc_fwd_tc" else 8) and s[k]["regs"] <= 128, (k, dict(s[k]))
    fwd = [k for k in s if k.split("<")[0] == "k_enc_fwd_tc"]
    assert len(fwd) == 4 and all(s[k]["UTMASTG"] > 0 for k in fwd), fwd  # one / two views x single CTA / CTA pair
    assert all(d["HMMA"] == 0 for d in s.values())
    assert len(s) >= 35


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU step, oracle port) needs no GPU: ONE JSON line on stdout carrying the
    bench contract's keys for the reference arm; a non-zero rank under torchrun prints nothing and exits 0."""
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DAD train samples/sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[3]" in d["config"]["workload"] and d["vs_baseline"] is None
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_bench_dominant_kernel_is_averaged_over_its_launches_per_step():
    """bench.py's roofline entry: the dominant kernel SYMBOL (k_enc_fwd_tc: the clean launch and the fused weak+strong launch of a step;
    k_enc_bwd_tc: clean and strong) with its AVERAGE launch over those variants -- not the fastest variant -- achieved bytes / flops
    summed over the variants, and the ncu DRAM traffic averaged the same way."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    v = lambda sym, ms, by, fl, pr: {"symbol": sym, "ms": ms, "alg_bytes": by, "alg_flops": fl, "products": pr}
    kern = {"encoder_fwd_clean": v("k_enc_fwd_tc", 0.9, 2.0e9, 4.0e11, 3), "encoder_fwd_noisy_fused": v("k_enc_fwd_tc", 2.1, 2.0e9, 8.0e11, 3),
            "encoder_bwd_clean": v("k_enc_bwd_tc", 1.2, 2.0e9, 4.0e11, 2), "encoder_bwd_strong": v("k_enc_bwd_tc", 1.0, 2.0e9, 4.0e11, 2)}
    table = {"encoder_fwd_clean": 10.0, "encoder_fwd_noisy_fused": 30.0, "encoder_bwd_clean": 1.0, "encoder_bwd_strong": 3.0, "_source": "x"}
    d = bench.dominant_kernel(kern, table)
    assert d["kernel"] == "k_enc_fwd_tc" and d["launches_per_step"] == 2 and abs(d["avg_launch_ms"] - 1.5) < 1e-12
    assert abs(d["GBps"] - 4.0e9 / 3.0 / 1e6) < 1e-6 and abs(d["TFLOPs"] - 12.0e11 / 3.0 / 1e9) < 1e-6 and d["traffic"] == 20.0
    assert abs(d["tensor_TFLOPs"] - 3 * d["TFLOPs"]) < 1e-9 and d["alg_bytes_per_launch"] == 2.0e9
    kern["encoder_bwd_clean"]["ms"] = 3.0
    d = bench.dominant_kernel(kern, {"encoder_bwd_clean": 1.0})
    assert d["kernel"] == "k_enc_bwd_tc" and d["launches_per_step"] == 2 and d["avg_launch_ms"] == 2.0 and d["traffic"] is None
    assert abs(d["tensor_TFLOPs"] - 2 * d["TFLOPs"]) < 1e-9
