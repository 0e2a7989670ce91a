#!/bin/bash
# Round-end evidence on 8 B200s of one box (gpurun --gpus 8): the headline and the fused workload under torchrun, the 2-GPU equality test
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
o=gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $o/r02_bench_${N}gpu.json 2> $o/r02_bench_${N}gpu.err
$T --master-port 29512 bench.py --gpus $N --workload fused --steps 5 --warmup 3 > $o/r02_bench_fused_${N}gpu.json 2> $o/r02_bench_fused_${N}gpu.err
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q > $o/r02_pytest_distributed.txt 2>&1
tail -2 $o/r02_pytest_distributed.txt
for f in r02_bench_${N}gpu r02_bench_fused_${N}gpu; do python - <<P
import json
try:
    d=json.loads(open("$o/$f.json").read().strip().splitlines()[-1]); print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("strong_scaling"), d.get("sharded_equals_single"), (d.get("lod_grid") or {}).get("grid",{}).get("wall_s"))
except Exception as e: print("$f", "ERR", e)
P
done
