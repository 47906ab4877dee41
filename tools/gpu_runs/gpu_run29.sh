set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default ss6 ss7 tn5 tn5w sa5 ss6sa5
cp gpurun_out/sweep.txt gpurun_out/r29_sweep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r29_pytest.txt
cat gpurun_out/r29_pytest.txt
