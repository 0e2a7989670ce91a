#!/bin/bash
# Round-end evidence on ONE B200: pytest -m gpu, every bench workload + the reference arms + CLI wall times -> gpurun_out/r02_*
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $o/r02_pytest.txt 2>&1
python bench.py --steps 5 --warmup 3 > $o/r02_bench_1gpu.json 2> $o/r02_bench_1gpu.err
python bench.py --workload fused --steps 5 --warmup 3 > $o/r02_bench_fused_1gpu.json 2> $o/r02_bench_fused_1gpu.err
python bench.py --workload external --steps 5 --warmup 3 > $o/r02_bench_external_1gpu.json 2> $o/r02_bench_external_1gpu.err
python bench.py --workload panel --steps 5 --warmup 3 > $o/r02_bench_panel_1gpu.json 2> $o/r02_bench_panel_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > $o/r02_bench_reference_arm.json 2> $o/r02_ref.err
python bench.py --impl reference --workload external --steps 2 --warmup 1 > $o/r02_bench_reference_arm_external.json 2> $o/r02_ref_ext.err
python tools/bench_collapse.py > $o/r02_bench_collapse_general.json 2> $o/r02_bc.err
bash tools/cli_walltimes.sh $o/r02_cli_walltimes.txt > /dev/null 2>&1
PMRD_SORT_BITS=24 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lod > $o/r02_bench_1gpu_sort24.json 2> /dev/null
tail -2 $o/r02_pytest.txt
for f in r02_bench_1gpu r02_bench_fused_1gpu r02_bench_external_1gpu r02_bench_panel_1gpu r02_bench_reference_arm r02_bench_reference_arm_external r02_bench_1gpu_sort24; do python - <<P
import json
try:
    d=json.loads(open("$o/$f.json").read().strip().splitlines()[-1]); print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"))
except Exception as e: print("$f", "ERR", e)
P
done
