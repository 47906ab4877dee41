#!/bin/bash
# Run on the GPU box: for every prebuilt variant of the library (reina-vk_b200/csrc/variants/<name>.so, built here by
# tools/build_variants.sh — nvcc cross-compiles without a GPU, so no GPU minutes are spent compiling) run the traversal
# micro-benchmark and a short bench, one summary line each, into gpurun_out/sweep.txt.
# usage: tools/sweep_variants.sh [name ...]        (default: every variant + the default build)
#   SWEEP_BENCH=0 skips bench.py; SWEEP_ENV="RB200_LANES=4 RB200_ENGINES=1" is applied to every run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/sweep.txt
names=("$@")
if [ ${#names[@]} -eq 0 ]; then names=(default $(ls reina-vk_b200/csrc/variants/*.so 2>/dev/null | xargs -n1 basename | sed 's/\.so$//')); fi
for n in "${names[@]}"; do
  lib=""; [ "$n" != default ] && lib="$PWD/reina-vk_b200/csrc/variants/$n.so"
  echo "== $n $SWEEP_ENV" | tee -a $out
  if [ "${SWEEP_TRAV:-1}" != 0 ]; then env $SWEEP_ENV RB200_LIBRARY=$lib timeout 300 python tools/trav_bench.py 2>&1 | tail -1 | tee -a $out; fi
  if [ "${SWEEP_BENCH:-1}" != 0 ]; then
    env $SWEEP_ENV RB200_LIBRARY=$lib timeout 600 python bench.py --steps ${SWEEP_STEPS:-8} --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; k=r['kernel_ms_per_batch']
    print('  total %.1f Mrays/s  step %.2f ms  e2e %.1f | frac %.3f | extend %.1f Mrays/s (%.2f ms, %d launches) shadow %.1f Mrays/s (%.2f ms) disney %.2f lamb %.2f miss %.2f finish %.2f | nodes/ray %.2f tris/ray %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['extend_mrays_s'], k['extend'], r['extend_launches'], r['shadow_mrays_s'], k['shadow'], k['shade_disney'], k['shade_lambertian'], k['miss'], k['finish'], r['nodes_per_ray'], r['tris_per_ray']))
except Exception as e:
    print('  bench failed:', e)
" | tee -a $out
  fi
done
