set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 tools/sweep_variants.sh default fu2 bl fu2bl default
cp gpurun_out/sweep.txt gpurun_out/r49_sweep.txt
