set -x
mkdir -p gpurun_out
M="gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 ncu --metrics $M --clock-control none -k regex:'k_extend|k_shadow' -s 700 -c 4 --csv --log-file gpurun_out/r5_$name.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r5_$name.log 2>&1
}
run default RB200_ENGINES=1 RB200_LANES=4
run evl RB200_ENGINES=1 RB200_LANES=4 RB200_LIBRARY=$PWD/reina-vk_b200/csrc/variants/evl.so
run persist RB200_ENGINES=1 RB200_LANES=4 RB200_L2_PERSIST=1
run evl_persist RB200_ENGINES=1 RB200_LANES=4 RB200_L2_PERSIST=1 RB200_LIBRARY=$PWD/reina-vk_b200/csrc/variants/evl.so
SWEEP_ENV="RB200_ENGINES=2 RB200_LANES=4" SWEEP_STEPS=12 tools/sweep_variants.sh default binm
SWEEP_ENV="RB200_ENGINES=2 RB200_LANES=4 RB200_L2_PERSIST=1" SWEEP_STEPS=12 tools/sweep_variants.sh default
cp gpurun_out/sweep.txt gpurun_out/r5_sweep.txt
