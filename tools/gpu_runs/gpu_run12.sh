set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=16 tools/sweep_variants.sh default ofs fmad
for cfg in "2 12" "2 16" "3 8" "1 16"; do set -- $cfg; SWEEP_ENV="RB200_ENGINES=$1 RB200_LANES=$2" SWEEP_STEPS=32 tools/sweep_variants.sh default; done
cp gpurun_out/sweep.txt gpurun_out/r12_sweep.txt
