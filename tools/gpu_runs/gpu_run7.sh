set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r7_gpus.txt
timeout 600 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -15 > gpurun_out/r7_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r7_bench2.log 2>&1
tail -3 gpurun_out/r7_bench2.log
timeout 300 ./reina-vk_b200/host/reina_b200 --config config/config.toml --gpus 2 --spp 32 --width 480 --height 270 --out gpurun_out/r7_cli2.png > gpurun_out/r7_cli2.log 2>&1
timeout 300 ./reina-vk_b200/host/reina_b200 --config config/config.toml --gpus 2 --tiles --spp 16 --width 480 --height 270 --out gpurun_out/r7_cli2t.png > gpurun_out/r7_cli2t.log 2>&1
timeout 300 ./reina-vk_b200/host/reina_b200 --config config/config.toml --spp 16 --width 480 --height 270 --out gpurun_out/r7_cli1.png > gpurun_out/r7_cli1.log 2>&1
cat gpurun_out/r7_cli2.log gpurun_out/r7_cli2t.log gpurun_out/r7_cli1.log
cmp gpurun_out/r7_cli2t.png gpurun_out/r7_cli1.png && echo "tiles image == single-GPU image"
