# ncu --set full of one steady-state launch of every wave kernel, build the round ends with (source tables are made from it)
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-as-shipped"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade|k_finish' -s 3300 -c 7 -o gpurun_out/r02z_wave $B > gpurun_out/r02z_ncu.log 2>&1
tail -2 gpurun_out/r02z_ncu.log | cut -c1-200
