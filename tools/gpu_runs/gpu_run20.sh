set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_two_level.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r20_pytest.txt
cat gpurun_out/r20_pytest.txt
timeout 600 python tools/two_level_bench.py > gpurun_out/r02i_two_level_bench.jsonl 2> gpurun_out/r02i_two_level_bench.err
cat gpurun_out/r02i_two_level_bench.jsonl; tail -3 gpurun_out/r02i_two_level_bench.err
