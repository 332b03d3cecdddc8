"""Build libdad_sm100a.so in-tree: nvcc -gencode arch=compute_100a,code=sm_100a, one object per .cu, linked shared.
Usage: python build.py [--force] [--verbose]   (also called by __graft_entry__.build())."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
INC = os.path.join(ROOT, "include")
OUT = os.path.join(os.path.dirname(HERE), "libdad_sm100a.so")
SRCS = ["dad_abi.cu", "k_aug.cu", "k_enc_simt.cu", "k_enc_tc.cu", "k_head.cu", "k_ecda.cu", "k_gram_tc.cu", "k_scl.cu", "k_peer.cu", "k_optim.cu", "k_predict.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-I", INC, "-I", HERE]


def _newer(a, deps):
    if not os.path.exists(a):
        return False
    t = os.path.getmtime(a)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))] + [os.path.join(INC, "dad_sm100a.h")]
    jobs = []
    for s in SRCS:
        src, obj = os.path.join(HERE, s), os.path.join(objdir, s.replace(".cu", ".o"))
        if force or not _newer(obj, [src] + hdrs):
            jobs.append((src, obj))

    extra = os.environ.get("DAD_NVCC_FLAGS", "").split()  # e.g. -DDAD_TC_EXPERIMENTS (timing-experiment switches, tools/gpu_dbg.sh)

    def cc(job):
        src, obj = job
        cmd = [nvcc] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for src, r in ex.map(cc, jobs):
            if verbose or r.returncode:
                sys.stderr.write(f"--- {os.path.basename(src)}\n{r.stdout}{r.stderr}\n")
            if r.returncode:
                raise RuntimeError(f"nvcc failed for {src}")
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SRCS]
    if force or jobs or not os.path.exists(OUT):
        r = subprocess.run([nvcc, "-shared", "-o", OUT] + objs + ["-lcuda"], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
