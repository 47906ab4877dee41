set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r14_pytest.txt
cat gpurun_out/r14_pytest.txt
SWEEP_STEPS=24 tools/sweep_variants.sh default chunk5 chunk8 refill24 ploc32 prim02 prim04 leaf4 stack6 blk128
cp gpurun_out/sweep.txt gpurun_out/r14_sweep.txt
