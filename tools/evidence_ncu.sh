#!/bin/bash
# Round-end ncu evidence on ONE B200 (each capture only after the same command exited 0 without ncu) -> gpurun_out/r02n_*
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
o=gpurun_out
export PMRD_BENCH_REPS=100
C="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-lod"
$C > $o/r02n_plain.json 2> $o/r02n_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $o/r02_launches.csv $C > $o/r02n_l.log 2>&1
# one whole step of the chunk: skip the plan build + 3 warm-up steps' launches (found from the launch list: the last size kernel)
ncu --set full --clock-control none --import-source on -k regex:"reads_size|reads_gen|rs_downsweep_staged|rs_downsweep_direct|rs_upsweep|rs_cellscan|cell_consensus" --launch-skip 24 -c 8 -f -o $o/r02n_pipeline $C > $o/r02n_p.log 2>&1
python bench.py --workload fused --steps 1 --warmup 3 --no-cpu-baseline --no-lod > $o/r02n_fplain.json 2> $o/r02n_fplain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:fused_cell_kernel --launch-skip 2 -c 1 -f -o $o/r02n_fused python bench.py --workload fused --steps 1 --warmup 3 --no-cpu-baseline --no-lod > $o/r02n_f.log 2>&1
unset PMRD_BENCH_REPS
python bench.py --workload external --steps 1 --warmup 3 --no-cpu-baseline > $o/r02n_eplain.json 2> $o/r02n_eplain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"group_rows|rs_downsweep_kernel|rs_upsweep|grp_bounds|grp_max|rs_varying" --launch-skip 20 -c 8 -f -o $o/r02n_external python bench.py --workload external --steps 1 --warmup 3 --no-cpu-baseline > $o/r02n_e.log 2>&1
ls -la $o/r02n_*.ncu-rep
