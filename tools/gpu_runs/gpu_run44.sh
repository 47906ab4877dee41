# eight GPUs, final build: latency mode with peer stores against the reduce; sample split on C3 and C5
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r44_gpus.txt
timeout 300 python tools/tile_latency.py --gpus 4,8 --config c3 --frames 16 > gpurun_out/r44_tile_latency_peer.jsonl 2> gpurun_out/r44_tile_latency_peer.err
cut -c1-420 gpurun_out/r44_tile_latency_peer.jsonl
RB200_GROUP_TILES_REDUCE=1 timeout 300 python tools/tile_latency.py --gpus 4,8 --config c3 --frames 16 > gpurun_out/r44_tile_latency_reduce.jsonl 2> gpurun_out/r44_tile_latency_reduce.err
cut -c1-420 gpurun_out/r44_tile_latency_reduce.jsonl
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-as-shipped > gpurun_out/r44_bench_8gpu_c3.log 2>&1
tail -1 gpurun_out/r44_bench_8gpu_c3.log | cut -c1-300
timeout 400 $T --master-port 29522 bench.py --gpus 8 --steps 12 --warmup 3 --config c5 --no-cpu-baseline --no-as-shipped --no-roofline > gpurun_out/r44_bench_8gpu_c5.log 2>&1
tail -1 gpurun_out/r44_bench_8gpu_c5.log | cut -c1-300
