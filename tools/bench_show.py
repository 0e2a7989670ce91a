import json,sys
for f in sys.argv[1:]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"]); print({k:v for k,v in d["kernels_ms_per_step"].items() if v>1})
