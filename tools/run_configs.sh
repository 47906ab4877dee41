#!/bin/bash
# Run on the GPU box: bench.py for every BASELINE config (one GPU), one JSON line each into gpurun_out/configs/<c>.json,
# then tools/collect_configs.py folds them into profiles/r02_configs.json (run here afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/configs
for c in c1 c2 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps ${CFG_STEPS:-16} --warmup 3 --no-cpu-baseline > gpurun_out/configs/$c.json 2> gpurun_out/configs/$c.err
  tail -c 400 gpurun_out/configs/$c.json; echo
done
