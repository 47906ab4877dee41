set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1_pytest.txt
rm -f gpurun_out/sweep.txt
for cfg in "1 4" "2 4" "4 1" "1 6" "2 2"; do set -- $cfg; SWEEP_ENV="RB200_ENGINES=$1 RB200_LANES=$2" SWEEP_STEPS=8 tools/sweep_variants.sh default; done
tools/sweep_variants.sh evl l1a l1b shcs evl_shcs
cp gpurun_out/sweep.txt gpurun_out/r1_sweep.txt
