set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh pc0 default pc0 default
cp gpurun_out/sweep.txt gpurun_out/r42_sweep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r42_pytest.txt
cat gpurun_out/r42_pytest.txt
