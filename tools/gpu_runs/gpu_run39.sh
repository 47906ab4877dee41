# FINAL records of the round for the build with six / five shared stack entries (profiles/r02z_*)
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
tail -c 600 gpurun_out/r02z_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_reference_arm.json 2> gpurun_out/r02z_reference_arm.err
tail -c 600 gpurun_out/r02z_reference_arm.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-as-shipped"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 700 --csv --log-file gpurun_out/r02z_launches.csv $B > gpurun_out/r02z_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade|k_finish' -s 3300 -c 7 -o gpurun_out/r02z_wave $B > gpurun_out/r02z_ncu.log 2>&1
CFG_STEPS=16 bash tools/run_configs.sh
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02z_sustained_clocks.csv &
SMI=$!
python bench.py --steps 128 --warmup 3 --no-cpu-baseline --no-roofline --no-as-shipped > gpurun_out/r02z_sustained.json 2> gpurun_out/r02z_sustained.err
kill $SMI
tail -c 400 gpurun_out/r02z_sustained.json
