set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
for cfg in "1 4" "2 4" "2 3" "2 6" "3 4" "1 8"; do set -- $cfg; SWEEP_ENV="RB200_ENGINES=$1 RB200_LANES=$2" SWEEP_STEPS=12 tools/sweep_variants.sh default; done
SWEEP_STEPS=12 tools/sweep_variants.sh h2 cp20 cp45 cp60 cp90
cp gpurun_out/sweep.txt gpurun_out/r3_sweep.txt
