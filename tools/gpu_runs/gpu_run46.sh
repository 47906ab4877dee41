# RB200_FLAG_SKIP_NULL_SHADOW_RAYS in k_shadow: parity, then the bench with its extra arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r46_pytest.txt
cat gpurun_out/r46_pytest.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r46_bench.json 2> gpurun_out/r46_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r46_bench.json')); print(d['value'], d['ms_per_step'], d['spp_per_s'], d['roofline']['kernel_ms_per_batch']['shadow']); print(json.dumps(d['null_shadow_rays_skipped']))"
