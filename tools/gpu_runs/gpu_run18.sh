set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r18_gpus.txt
timeout 400 python tools/tile_latency.py --gpus 1,2,4,8 --config c3 --frames 12 > gpurun_out/r02g_tile_latency.jsonl 2> gpurun_out/r02g_tile_latency.err
cat gpurun_out/r02g_tile_latency.jsonl; tail -3 gpurun_out/r02g_tile_latency.err
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-as-shipped > gpurun_out/r02g_bench_8gpu_c3.log 2>&1
tail -1 gpurun_out/r02g_bench_8gpu_c3.log | cut -c1-600
timeout 500 $T --master-port 29522 bench.py --gpus 8 --steps 12 --warmup 3 --config c5 --no-cpu-baseline --no-as-shipped --no-roofline > gpurun_out/r02g_bench_8gpu_c5.log 2>&1
tail -1 gpurun_out/r02g_bench_8gpu_c5.log | cut -c1-600
