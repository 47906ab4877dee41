set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r9_pytest.txt
cat gpurun_out/r9_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; tail -c 3000 gpurun_out/r9_bench.json
