"""Summarise an .ncu-rep (read here, no GPU): key raw metrics per kernel + top stall sites of the SASS source page."""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum "]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pat and pat not in d["Kernel Name"]:
        continue
    print("====", d["Kernel Name"][:90])
    for k in hdr:
        if any((k + " ").startswith(x) or k == x.strip() for x in KEYS):
            print(f"   {k:85s} {d[k]:>16s} {units[hdr.index(k)]}")
# source page: one CSV with a section per captured launch ("Kernel Name",<name> row, then the header row starting with "Address")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for sec in secs:
    kn = sec["name"]
    if (pat and pat not in kn) or not sec["rows"] or sec["rows"][0][0] != "Address":
        continue
    hd = sec["rows"][0]; ix = {k: i for i, k in enumerate(hd)}
    data = [r for r in sec["rows"][1:] if len(r) == len(hd) and r[0].startswith("0x")]
    key = (kn, sum(int(r[ix["# Samples"]]) for r in data))
    if key in seen:  # ncu repeats a section per source view
        continue
    seen.add(key)
    short = kn.split("(CUtensorMap")[0].split("::")[-1][:60]
    stalls = [k for k in hd if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    ex = sum(int(r[ix["Instructions Executed"]]) for r in data)
    agg = {k: sum(int(r[ix[k]]) for r in data) for k in stalls}
    print(f"---- {short}: {tot} samples, {ex} warp-instructions executed (source-counter pass); stalls:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
    op = collections.Counter()
    for r in data:
        t = r[ix["Source"]].split()
        if not t: continue
        o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        op[o.split(".")[0]] += int(r[ix["Instructions Executed"]])
    print("     opcode mix:", ", ".join(f"{k}={100*v/max(ex,1):.1f}%" for k, v in op.most_common(14)))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:12]
    for i in sorted(top):
        r = data[i]
        st = {k[6:]: int(r[ix[k]]) for k in stalls if int(r[ix[k]]) > 0}
        print(f"     [{i:5d}] {int(r[ix['# Samples']]):6d}  {r[ix['Source']].strip()[:60]:60s} {st}")
