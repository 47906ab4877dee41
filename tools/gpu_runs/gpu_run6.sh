set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r6_pytest.txt
for cfg in "2 4" "2 6" "2 8" "3 4" "3 6" "4 4" "4 6"; do set -- $cfg; SWEEP_ENV="RB200_ENGINES=$1 RB200_LANES=$2" SWEEP_STEPS=24 tools/sweep_variants.sh default; done
cp gpurun_out/sweep.txt gpurun_out/r6_sweep.txt
