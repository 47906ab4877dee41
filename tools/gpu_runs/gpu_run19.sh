set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default dmb5 dmb6 dmb3
cp gpurun_out/sweep.txt gpurun_out/r19_sweep.txt
