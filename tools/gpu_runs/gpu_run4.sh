set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r4_pytest.txt
export SWEEP_ENV="RB200_ENGINES=2 RB200_LANES=4" SWEEP_STEPS=12
tools/sweep_variants.sh default de0 ns6 ns8 ch8 ch4 wc128
for c in 50 63 72 86; do SWEEP_ENV="RB200_ENGINES=2 RB200_LANES=4 RB200_SMEM_CARVEOUT=$c" tools/sweep_variants.sh default; done
SWEEP_ENV="RB200_ENGINES=2 RB200_LANES=8" tools/sweep_variants.sh default
SWEEP_ENV="RB200_ENGINES=3 RB200_LANES=6" tools/sweep_variants.sh default
cp gpurun_out/sweep.txt gpurun_out/r4_sweep.txt
