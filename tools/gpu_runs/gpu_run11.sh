set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-as-shipped"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 700 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade|k_finish' -s 3300 -c 7 -o gpurun_out/r02b_wave $B > gpurun_out/r02b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_blur|k_combine' -s 9 -c 3 -o gpurun_out/r02b_post python tools/post_profile.py c3 > gpurun_out/r02b_post.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02_sustained_clocks.csv &
SMI=$!
python bench.py --steps 128 --warmup 3 --no-cpu-baseline --no-roofline --no-as-shipped > gpurun_out/r02_sustained.json 2> gpurun_out/r02_sustained.err
kill $SMI
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 600 gpurun_out/r02_sustained.json
ls -la gpurun_out | tail -12
