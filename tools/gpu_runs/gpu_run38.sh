set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default sm6 sm7 dm7 sm6dm7
cp gpurun_out/sweep.txt gpurun_out/r38_sweep.txt
