# two GPUs: latency mode with peer stores against the reduce
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r43_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -15 > gpurun_out/r43_pytest.txt
cat gpurun_out/r43_pytest.txt
timeout 400 python tools/tile_latency.py --gpus 1,2 --config c3 --frames 16 > gpurun_out/r43_tile_latency_peer.jsonl 2> gpurun_out/r43_tile_latency_peer.err
cut -c1-420 gpurun_out/r43_tile_latency_peer.jsonl
RB200_GROUP_TILES_REDUCE=1 timeout 400 python tools/tile_latency.py --gpus 2 --config c3 --frames 16 > gpurun_out/r43_tile_latency_reduce.jsonl 2> gpurun_out/r43_tile_latency_reduce.err
cut -c1-420 gpurun_out/r43_tile_latency_reduce.jsonl
timeout 300 ./reina-vk_b200/host/reina_b200 --config config/config.toml --gpus 2 --tiles --spp 16 --width 480 --height 270 --out gpurun_out/r43_cli2t.png > gpurun_out/r43_cli2t.log 2>&1
timeout 300 ./reina-vk_b200/host/reina_b200 --config config/config.toml --spp 16 --width 480 --height 270 --out gpurun_out/r43_cli1.png > gpurun_out/r43_cli1.log 2>&1
cmp gpurun_out/r43_cli2t.png gpurun_out/r43_cli1.png && echo "tiles image == single-GPU image"
