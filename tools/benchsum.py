"""One-line summary of bench.py JSON lines read from stdin (scratch helper for the profiling runs)."""
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    r=d.get("roofline") or {}
    k=r.get("kernels",{})
    print("value",round(d["value"],3),"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],2),"cg_it_ms",k.get("cg_iteration",{}).get("ms_per_launch"),"cg/step",r.get("cg_iterations_per_step"),"launches",d.get("gpu_launches"),"vcycle",k.get("vcycle_ms"))
