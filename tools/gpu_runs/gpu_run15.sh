set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 tools/sweep_variants.sh default pf0 pfmb5 pfmb6 ca3 ca4 ca5
for cfg in "RB200_TRAV_BLOCKS_PER_SM=3" "RB200_TRAV_BLOCKS_PER_SM=3 RB200_ENGINES=3" "RB200_TRAV_BLOCKS_PER_SM=3 RB200_SHADE_BLOCKS_PER_SM=2" "RB200_TRAV_BLOCKS_PER_SM=2 RB200_ENGINES=4 RB200_LANES=4"; do
  SWEEP_ENV="$cfg" SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default
done
cp gpurun_out/sweep.txt gpurun_out/r15_sweep.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r15_pytest.txt
cat gpurun_out/r15_pytest.txt
