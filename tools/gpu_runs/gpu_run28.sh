set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default frcp0 ss5 ss6 sa5 sa6
SWEEP_STEPS=24 SWEEP_TRAV=0 SWEEP_ENV="RB200_SMEM_CARVEOUT=100" tools/sweep_variants.sh default
cp gpurun_out/sweep.txt gpurun_out/r28_sweep.txt
