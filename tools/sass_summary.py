"""Per-kernel SASS evidence of the built library (no GPU needed): tensor-core / TMA / TMEM mnemonics, registers, local memory.
Usage: python tools/sass_summary.py [libdad_sm100a.so]    (cuobjdump -sass / -res-usage; B200_PROFILING.md "SASS mnemonics")

    UTCHMMA  = tcgen05.mma          UTCBAR  = tcgen05.commit -> mbarrier     LDTM = tcgen05.ld (TMEM -> registers)
    UTMALDG  = TMA tensor load      UTMASTG = TMA tensor store               SYNCS.* = mbarrier arrive / try_wait"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "robust-speech-emotion-recognition-via-dynamic-asymmetric-distillation-in-noisy-environments_b200", "libdad_sm100a.so")
MNEMONICS = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FFMA", "MUFU", "LDL", "STL")


_names = {}


def demangle_short(name):
    """_ZN..k_enc_fwd_tcE... -> k_enc_fwd_tc (template arguments kept: k_bwd_db<8>)"""
    if name not in _names:
        full = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        m = re.search(r"(k_[a-z0-9_]+(?:<[^>]*>)?)\(", full)
        _names[name] = m.group(1) if m else full
    return _names[name]


def summary(lib=LIB):
    """-> {kernel: {'regs': int, 'stack': bytes of local memory per thread, 'smem': static bytes, mnemonic: count, ...}}"""
    cuobjdump = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "cuobjdump")
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run([cuobjdump, "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    out = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = out.setdefault(demangle_short(m.group(1)), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            cur["instructions"] += 1
            if op in MNEMONICS:
                cur[op] += 1
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+?):", line)
        if m:
            fn = demangle_short(m.group(1))
            continue
        if fn and "REG:" in line:
            d = out.setdefault(fn, collections.Counter())
            for key, field in (("regs", "REG"), ("stack", "STACK"), ("smem", "SHARED")):
                m = re.search(field + r":(\d+)", line)
                if m:
                    d[key] = max(d[key], int(m.group(1)))
            fn = None
    return out


if __name__ == "__main__":
    s = summary(sys.argv[1] if len(sys.argv) > 1 else LIB)
    cols = ("instructions", "regs", "stack", "smem") + MNEMONICS
    print(f"{'kernel':28s}" + "".join(f"{c:>9s}" for c in cols))
    for k, d in s.items():
        print(f"{k[:28]:28s}" + "".join(f"{d.get(c, 0):9d}" for c in cols))
