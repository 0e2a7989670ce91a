#!/bin/bash
# eval-lod --engine grid from the CLI on 1 GPU and under torchrun on N GPUs: the artifacts must be identical; wall times printed.
# usage (on a box with N GPUs, from the repo root): tools/cli_multigpu_check.sh N OUT.txt
N=${1:-2}; out=${2:-gpurun_out/cli_multigpu.txt}
root=$(pwd); export PYTHONPATH="$root/precise-mrd-mini_b200"
a=$(mktemp -d); b=$(mktemp -d)
: > "$out"
( cd "$a"; python -m precise_mrd.cli --seed 7 --engine grid eval-lod --replicates 25 > /dev/null 2>&1 )   # warm the page cache
t0=$(date +%s%N); ( cd "$a"; python -m precise_mrd.cli --seed 7 --engine grid eval-lod --replicates 25 > run.log 2>&1 ); rc1=$?; t1=$(date +%s%N)
( cd "$b"; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 -m precise_mrd.cli --seed 7 --engine grid eval-lod --replicates 25 > run.log 2>&1 ); rc2=$?; t2=$(date +%s%N)
echo "eval-lod --engine grid --replicates 25: 1 GPU rc=$rc1 $(( (t1 - t0) / 1000000 )) ms wall; $N GPUs (torchrun) rc=$rc2 $(( (t2 - t1) / 1000000 )) ms wall" >> "$out"
if cmp -s "$a/reports/lod_table.csv" "$b/reports/lod_table.csv" && cmp -s "$a/reports/lob.json" "$b/reports/lob.json"; then echo "lod_table.csv and lob.json identical" >> "$out"; else echo "ARTIFACTS DIFFER" >> "$out"; tail -5 "$b/run.log" >> "$out"; fi
cat "$a/reports/lod_table.csv" >> "$out"
rm -rf "$a" "$b"; cat "$out"
