set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r22_pytest.txt
cat gpurun_out/r22_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
tail -c 600 gpurun_out/r02k_bench.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r22_smoke.txt 2>&1; tail -2 gpurun_out/r22_smoke.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02k_reference.json 2> gpurun_out/r02k_reference.err; tail -c 400 gpurun_out/r02k_reference.json
