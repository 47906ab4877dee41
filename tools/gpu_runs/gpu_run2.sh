set -x
mkdir -p gpurun_out
# steady-state capture of one wave's kernels (engine 1 x 4): skip the two cold calls (2 x 128 waves x 5 matching kernels)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 1700 -c 10 \
    -o gpurun_out/r02a_wave python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
ls -la gpurun_out/
