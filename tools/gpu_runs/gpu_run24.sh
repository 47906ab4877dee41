set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default tb3 b128m7 ch5 ch7 ca3 ca5
cp gpurun_out/sweep.txt gpurun_out/r24_sweep.txt
