"""Summarise an .ncu-rep (read here, no GPU needed) into a markdown table for profiles/."""
import csv
import subprocess
import sys
from collections import OrderedDict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = OrderedDict([
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %")])
print("| kernel | " + " | ".join(cols.values()) + " | top stalls |")
print("|---|" + "---|" * (len(cols) + 1))
seen = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("ddps::", "")
    if "<" in name:
        short += name[name.find("<"):name.find(">") + 1].replace("(int)", "").replace("(bool)", "")
    seen[short] = seen.get(short, 0) + 1
    if seen[short] > 1:
        continue
    vals = []
    for c in cols:
        if c in ix and r[ix[c]] not in ("", "n/a"):
            try:
                v = float(r[ix[c]].replace(",", ""))
                u = units[ix[c]]
                vals.append(f"{v:.4g} {u}".strip())
            except ValueError:
                vals.append(r[ix[c]])
        else:
            vals.append("-")
    st = []
    for h, i in ix.items():
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:3])
    print(f"| `{short}` | " + " | ".join(vals) + f" | {top} |")
