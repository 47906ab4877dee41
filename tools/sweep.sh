#!/bin/bash
# Run on the GPU box: rebuild librb200.so with each EXTRA flag set and print the traversal numbers of one bench run.
# usage: tools/sweep.sh "-DRB_REFILL=16" "-DRB_REFILL=28 -DRB_CHUNK=2" ...
cd "$(dirname "$0")/.."
for extra in "$@"; do
  make -C reina-vk_b200/csrc clean >/dev/null
  make -C reina-vk_b200/csrc -j8 EXTRA="$extra" >/dev/null 2>&1 || { echo "BUILD FAILED: $extra"; continue; }
  python bench.py --steps ${SWEEP_STEPS:-6} --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; k=r['kernel_ms_per_batch']
print('EXTRA=[$extra] total %.1f Mrays/s  step %.1f ms  e2e %.1f | extend %.1f Mrays/s (%.1f ms) shadow %.1f Mrays/s (%.1f ms) disney %.1f lamb %.1f finish %.1f | nodes/ray %.2f tris/ray %.2f wide nodes %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['extend_mrays_s'], k['extend'], r['shadow_mrays_s'], k['shadow'], k['shade_disney'], k['shade_lambertian'], k['finish'], r['nodes_per_ray'], r['tris_per_ray'], d['bvh']['numWideNodes']))"
done
make -C reina-vk_b200/csrc clean >/dev/null; make -C reina-vk_b200/csrc -j8 >/dev/null 2>&1
