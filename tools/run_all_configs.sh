#!/bin/bash
# Runs every bench configuration once (one JSON line each) into $1 (default gpurun_out/all_configs.jsonl).
out=${1:-gpurun_out/all_configs.jsonl}
steps=${2:-200}
: > "$out"
for c in headline cfg1 cfg2 ssd300 retina800 cfg4 crowd512; do
  python bench.py --config $c --steps $steps --warmup 10 >> "$out" 2>> "${out%.jsonl}.err" || echo "{\"config\": \"$c\", \"failed\": true}" >> "$out"
done
