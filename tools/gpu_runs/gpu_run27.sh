set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r27_pytest.txt
cat gpurun_out/r27_pytest.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default defer0 frcp dom frcpdom e8d e256d all8
cp gpurun_out/sweep.txt gpurun_out/r27_sweep.txt
