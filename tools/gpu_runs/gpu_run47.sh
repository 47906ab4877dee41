set -x
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
SWEEP_STEPS=24 SWEEP_TRAV=0 tools/sweep_variants.sh default default
cp gpurun_out/sweep.txt gpurun_out/r47_sweep.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r47_pytest.txt
cat gpurun_out/r47_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r47_bench.json 2> gpurun_out/r47_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r47_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_batch']); print(d['null_shadow_rays_skipped'])"
