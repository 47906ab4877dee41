# latency mode with 16 lanes per engine in the members of a tile group: tests, then frame latency
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -3
if [ "$N" -ge 8 ]; then G=4,8; else G=1,2; fi
timeout 400 python tools/tile_latency.py --gpus $G --config c3 --frames 16 > gpurun_out/r51_tile_latency_$N.jsonl 2> gpurun_out/r51_tile_latency_$N.err
cut -c1-420 gpurun_out/r51_tile_latency_$N.jsonl
