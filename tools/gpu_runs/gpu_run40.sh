# two GPUs, final build: the group API tests, the 2-rank bench with its parity check, latency mode on 1 and 2 devices
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r40_gpus.txt
timeout 600 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -5 > gpurun_out/r40_pytest.txt
cat gpurun_out/r40_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline --no-as-shipped > gpurun_out/r40_bench2.log 2>&1
tail -1 gpurun_out/r40_bench2.log | cut -c1-1500
timeout 400 python tools/tile_latency.py --gpus 1,2 --config c3 --frames 12 > gpurun_out/r40_tile_latency.jsonl 2> gpurun_out/r40_tile_latency.err
cut -c1-400 gpurun_out/r40_tile_latency.jsonl
