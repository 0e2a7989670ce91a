#!/bin/bash
# Wall time of every CLI command, interpreter start to artifacts written ("from CLI entry to
# lod_table.csv written").  Usage: tools/cli_walltimes.sh OUT.txt   (run from the repo root on a B200)
out=${1:-gpurun_out/cli_walltimes.txt}
export PYTHONPATH=precise-mrd-mini_b200
work=$(mktemp -d)
root=$(pwd)
: > "$out"
cd "$work"
run() {
  local label="$1"; shift
  local t0=$(date +%s%N)
  PYTHONPATH="$root/precise-mrd-mini_b200" python -m precise_mrd.cli "$@" > "$work/last.log" 2>&1
  local rc=$?
  local t1=$(date +%s%N)
  local ms=$(( (t1 - t0) / 1000000 ))
  printf "%-46s rc=%d  %d.%03d s\n" "$label" "$rc" $((ms / 1000)) $((ms % 1000)) >> "$root/$out"
}
run "warm-up (imports, CUDA context)"          --seed 7 smoke
run "smoke (stages engine)"                    --seed 7 smoke
run "eval-lob --n-blank 50 (grid engine)"      --seed 7 --engine grid eval-lob --n-blank 50
run "eval-lod --replicates 25 (grid engine)"   --seed 7 --engine grid eval-lod --replicates 25
run "eval-loq --replicates 25 (grid engine)"   --seed 7 --engine grid eval-loq --replicates 25
run "eval-contamination (grid engine)"         --seed 7 --engine grid eval-contamination
run "eval-lod --replicates 2 (stages engine)"  --seed 7 eval-lod --replicates 2
run "eval-stratified (stages engine)"          --seed 7 eval-stratified
ls "$work/reports" >> "$root/$out" 2>/dev/null
cd "$root"; rm -rf "$work"
cat "$out"
