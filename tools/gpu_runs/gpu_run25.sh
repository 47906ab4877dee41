set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default e7 e8 e96 s128
cp gpurun_out/sweep.txt gpurun_out/r25_sweep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r25_pytest.txt
cat gpurun_out/r25_pytest.txt
