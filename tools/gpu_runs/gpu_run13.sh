set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 tools/sweep_variants.sh default ofs fmad m16 m18 m0
cp gpurun_out/sweep.txt gpurun_out/r13_sweep.txt
