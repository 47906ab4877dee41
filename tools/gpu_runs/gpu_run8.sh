set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity2.py tests/test_spirv_hits.py tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r8_pytest.txt
cat gpurun_out/r8_pytest.txt
