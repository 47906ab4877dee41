set -x
mkdir -p gpurun_out
python tools/debug_surface_rays.py 2>&1 | head -4
timeout 600 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
bash tools/run_configs.sh
