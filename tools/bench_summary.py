import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["step_ms_rank0"], d["e2e"]["ms_per_step"], d["parity"]); print({k:(v["launches_per_step"],v["ms_per_step"]) for k,v in d["kernels"].items()})
