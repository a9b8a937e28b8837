"""Per-instruction stall summary of one kernel from an .ncu-rep (read here, no GPU): python tools/ncu_source.py rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines())]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address" or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break   # next kernel instance
        continue
    data.append(r)
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print("static", len(data), "warp insts executed", tot_inst, "samples", tot_samp)
for pos, r in enumerate(data):
    r.append(pos)
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
    st = {k: int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "Not" not in k and r[ix[k]] not in ("", "0")}
    st = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(str(r[-1]).rjust(5), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]].strip()[:64].ljust(64), st)
