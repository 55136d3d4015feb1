# round-2 ncu captures (one GPU; every command runs plain first, and only then under ncu)
#   sh tools/ncu_r02.sh [launches] [fwd_tc] [traffic] [fwd_tcl] [sgld]
set -x
mkdir -p gpurun_out
for what in "$@"; do
case $what in
launches)
  B1="python bench.py --steps 2 --warmup 1 --no-check --no-configs --sgld-steps 100"
  $B1 > gpurun_out/ncu_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $B1 > gpurun_out/ncu_l1.log 2>&1 ;;
fwd_tc)
  B2="python bench.py --steps 1 --warmup 1 --no-sgld --no-check --no-configs"
  $B2 > gpurun_out/ncu_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fwd_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02_fwd_tc $B2 > gpurun_out/ncu_l2.log 2>&1 ;;
traffic)
  B2="python bench.py --steps 1 --warmup 1 --no-sgld --no-check --no-configs"
  $B2 > gpurun_out/ncu_plain5.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:fwd_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_fwd_tc_traffic.csv $B2 > gpurun_out/ncu_l5.log 2>&1 ;;
fwd_tcl)
  B3="python tools/tcl_bench.py 50 100000"
  $B3 > gpurun_out/ncu_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fwd_tcl_kernel -s 1 -c 1 -f -o gpurun_out/r02_fwd_tcl $B3 > gpurun_out/ncu_l3.log 2>&1 ;;
sgld)
  B4="python tools/sgld_one.py 128"
  $B4 > gpurun_out/ncu_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sgld_step_kernel -s 1 -c 1 -f -o gpurun_out/r02_sgld_step $B4 > gpurun_out/ncu_l4.log 2>&1 ;;
esac
done
ls -la gpurun_out/*.ncu-rep
